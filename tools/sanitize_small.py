#!/usr/bin/env python3
"""Small runs of the persistent-warp and time-tiled kernels (ragged scenario counts, offsets, every output mode and RNG mode)
for compute-sanitizer (memcheck / racecheck).  usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import GBM_EQ, HESTON_EQ, grid  # noqa: E402

JUMP = ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
        "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
        "C = max(X1 - 100.0, 0.0) + X0"]
fast = dict(icdf="fast", arithmetic="fast")
runs = [
    ("resident C2 shape", GBM_EQ, grid(252, 37), {"X1": 1.0}, 333, 7, "euler", "sobol", dict(scramble="xor", **fast)),
    ("resident gamma/tails", GBM_EQ, grid(252, 10), {"X1": 1.0}, 131, 2, "euler", "sobol", dict(scramble="xor", **fast)),
    ("resident raw", GBM_EQ, grid(252, 9), {"X1": 1.0}, 65, 0, "euler", "sobol", dict(scramble="none")),
    ("resident heston euler", HESTON_EQ, grid(250, 21), {"S": 100.0, "v": 0.04}, 97, 5, "euler", "sobol", dict(scramble="xor", **fast)),
    ("tiled pseudo paths", GBM_EQ, grid(252, 19), {"X1": 1.0}, 301, 3, "euler", "pseudo", {}),
    ("tiled cp_shift", GBM_EQ, grid(252, 19), {"X1": 1.0}, 129, 1, "euler", "sobol", {}),
    ("tiled heston rk terminal", HESTON_EQ, grid(1000, 70), {"S": 100.0, "v": 0.04}, 300, 0, "runge-kutta", "sobol", dict(scramble="xor", output="terminal", **fast)),
    ("tiled moments", GBM_EQ, grid(365, 33), {"X1": 1.0}, 1000, 11, "euler", "pseudo", dict(output="moments", **fast)),
    ("tiled TPN", GBM_EQ, grid(252, 17), {"X1": 1.0}, 77, 0, "euler", "sobol", dict(scramble="xor", layout="TPN")),
    ("tiled jump rk", JUMP, grid(50, 15), {"X0": 0.4, "X1": 100.0}, 45, 0, "runge-kutta", "pseudo", {}),
    # round 2: four-buffer mbarrier tile hand-over (several tiles), factored RK step, Philox, bulk-copy and tensor-map stores
    ("tiled heston rk paths (mbarrier hand-over)", HESTON_EQ, grid(1000, 130), {"S": 100.0, "v": 0.04}, 300, 3, "runge-kutta", "sobol", dict(scramble="xor", **fast)),
    ("tiled heston rk paths, 1 tile + tail", HESTON_EQ, grid(1000, 23), {"S": 100.0, "v": 0.04}, 70, 0, "runge-kutta", "sobol", dict(scramble="xor", **fast)),
    ("tiled philox moments (hand-over)", GBM_EQ, grid(365, 140), {"X1": 1.0}, 700, 0, "euler", "pseudo", dict(output="moments", generator="philox", **fast)),
    ("tiled TPN fast (hand-over)", GBM_EQ, grid(252, 100), {"X1": 1.0}, 130, 2, "euler", "sobol", dict(scramble="xor", layout="TPN", **fast)),
    ("tiled f32", GBM_EQ, grid(252, 23), {"X1": 1.0}, 99, 4, "euler", "sobol", dict(scramble="xor", icdf="single", arithmetic="fast", dtype="f32")),
]
for name, eqs, times, init, N, off, scheme, rng, kw in runs:
    r = S.simulate(eqs, times, N, init, rng, scheme, seed=5, scenario_offset=off, **kw)
    print(f"{name:44s} {tuple(r.values.shape)} sum={float(r.values.double().sum()):.6g}", flush=True)
for mode, label in ((4, "bulk-copy stores"), (5, "tensor-map stores")):
    for off in (0, 251):                                  # with and without the negative-row warp
        plan = S.Plan(S.Universe(HESTON_EQ, grid(1000, 61)), "euler", "sobol", scramble="xor", ntp_direct=mode, **fast)
        v = plan.run({"S": 100.0, "v": 0.04}, 200, seed=5, scenario_offset=off)
        print(f"{label + ' offset ' + str(off):44s} {tuple(v.shape)} sum={float(v.double().sum()):.6g}", flush=True)
