"""The Rust side of the drop-in exists as source only (no toolchain in this image): hold it to the C header by text.
rust/sde-b200-sys/src/lib.rs must declare every function of include/sde_b200.h (and nothing else), and its `sde_options`
must list the header's fields in the header's order; the golden-vector harness must only use `pub` items of the reference."""
import os
import re

from conftest import ROOT

HEADER = open(os.path.join(ROOT, "include", "sde_b200.h")).read()
SYS = open(os.path.join(ROOT, "rust", "sde-b200-sys", "src", "lib.rs")).read()


def test_sys_crate_declares_exactly_the_header_functions():
    declared = set(re.findall(r"\b(sde_[a-z0-9_]+)\s*\(", HEADER)) - {"sde_b200"}
    bound = set(re.findall(r"pub fn (sde_[a-z0-9_]+)\s*\(", SYS))
    assert declared == bound, sorted(declared ^ bound)


def test_sys_crate_options_struct_matches_header_field_order():
    body = re.search(r"typedef struct sde_options \{(.*?)\} sde_options;", HEADER, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = re.findall(r"\b([a-z_0-9]+)\s*;", body)
    r_body = re.search(r"pub struct sde_options \{(.*?)\n\}", SYS, re.S).group(1)
    r_fields = re.findall(r"pub ([a-z_0-9]+):", r_body)
    assert c_fields == r_fields, (c_fields, r_fields)


def test_golden_harness_and_replacement_sources_exist():
    main = open(os.path.join(ROOT, "tools", "rust_golden", "src", "main.rs")).read()
    for item in ("PseudoRng::new", "SobolEngine::new", "SobolRng::new", "parse_equations", "euler_iteration", "runge_kutta_iteration",
                 "ScenarioFiltration::new", "WienerIncrementor::new", "PoissonJumpIncrementor::new", "Function::new"):
        assert item in main, item
    assert 'path = "/root/reference"' in open(os.path.join(ROOT, "tools", "rust_golden", "Cargo.toml")).read()
    sim = open(os.path.join(ROOT, "rust", "sde-sim-rs-b200", "sim_mod.rs")).read()
    assert "pub fn simulate(" in sim and "sde_simulate_devices" in sim and "PolarsResult<LazyFrame>" in sim
