#!/usr/bin/env python3
"""Lower a named config, write the generated translation unit to build/<name>.cu, compile it with nvcc for
sm_100a (-lineinfo, -Xptxas -v) and dump the SASS — the offline view of what NVRTC builds at plan creation."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sde-sim-rs_b200")
sys.path.insert(0, PKG)
import sde_sim_rs as S  # noqa: E402
from sde_sim_rs import _ffi  # noqa: E402

GBM = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
HESTON = ["dS = ( 0.05 * S ) * dt + ( max(v, 0.0)^0.5 * S ) * dW1",
          "dv = ( 2.0 * (0.04 - v) ) * dt + ( -0.21 * max(v, 0.0)^0.5 ) * dW1 + ( 0.2142428528562855 * max(v, 0.0)^0.5 ) * dW2"]
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import basket_equations  # noqa: E402

CONFIGS = {
    "c4": (basket_equations(64)[0], 252, "euler", "sobol", dict(scramble="xor", icdf="fast", arithmetic="fast", output="moments")),
    "c2": (GBM, 252, "euler", "sobol", dict(scramble="xor", icdf="fast", arithmetic="fast")),
    "c2strict": (GBM, 252, "euler", "sobol", dict(scramble="xor", icdf="reference", arithmetic="strict")),
    "c1": (GBM, 252, "euler", "pseudo", dict()),
    "c5": (GBM, 365, "euler", "pseudo", dict(output="moments", icdf="fast", arithmetic="fast")),
    "c5p": (GBM, 365, "euler", "pseudo", dict(output="moments", icdf="fast", arithmetic="fast", generator="philox")),
    "c2cp": (GBM, 252, "euler", "sobol", dict(icdf="fast", arithmetic="fast")),
    "c3t": (HESTON, 1000, "runge-kutta", "sobol", dict(scramble="xor", icdf="fast", arithmetic="fast", output="terminal")),
    "c3": (HESTON, 1000, "runge-kutta", "sobol", dict(scramble="xor", icdf="fast", arithmetic="fast")),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    eqs, D, scheme, rng, kw = CONFIGS[name]
    extra = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
    u = S.Universe(eqs, [k / D for k in range(D + 1)])
    o = S._make_options(device=0, seed=0, scenario_offset=0, output=kw.get("output", "paths"), layout=kw.get("layout", "NTP"),
                        scramble=kw.get("scramble", "cp_shift_per_path"), icdf=kw.get("icdf", "reference"),
                        arithmetic=kw.get("arithmetic", "strict"), rk_variant="reference",
                        tile_steps=int(extra.get("tt", 0)), block_threads=int(extra.get("block", 0)), ntp_direct=int(extra.get("direct", 0)), generator=kw.get("generator", "chacha8"))
    src = C.c_void_p()
    _ffi.check(_ffi.lib().sde_lower_only(u._h, scheme.encode(), rng.encode(), C.byref(o), 0, C.byref(src), None))
    text = C.string_at(src).decode()
    _ffi.lib().sde_free_string(src)
    bdir = os.path.join(PKG, "build")
    os.makedirs(bdir, exist_ok=True)
    cu, cubin, sass = (os.path.join(bdir, f"{name}.{e}") for e in ("cu", "cubin", "sass"))
    open(cu, "w").write(text)
    r = subprocess.run(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                        "-Xptxas", "-v", "-I", os.path.join(PKG, "csrc", "kernels"), "-o", cubin, cu], capture_output=True, text=True)
    print("\n".join(l for l in r.stderr.splitlines() if "registers" in l or "spill" in l or "error" in l))
    with open(sass, "w") as f:
        subprocess.run(["cuobjdump", "-sass", "-fun", "sde_sim_kernel", cubin], stdout=f)
    print(cu, sass)


if __name__ == "__main__":
    main()
