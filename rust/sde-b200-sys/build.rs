// Locates libsde_b200.so: $SDE_B200_LIB_DIR, else the in-tree build output of `python sde-sim-rs_b200/build.py`.
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=SDE_B200_LIB_DIR");
    let dir = env::var("SDE_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../sde-sim-rs_b200/sde_sim_rs")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=sde_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
