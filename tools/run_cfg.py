#!/usr/bin/env python3
"""Run one named BASELINE config a few times (device-resident output) and print its kernel time: the target of ncu
captures and A/B timings of everything that is not the C2 headline.
usage: run_cfg.py <c2|c3|c3t|c2cp|c2tpn|c5|c5p> [runs]      env SDE_B200_DEFINES etc. are honoured by the lowering"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import GBM_EQ, HESTON_EQ, grid  # noqa: E402

fast = dict(icdf="fast", arithmetic="fast")
CFG = {
    "c3": (HESTON_EQ, 1000, {"S": 100.0, "v": 0.04}, 1 << 22, "runge-kutta", "sobol", dict(output="paths", scramble="xor", **fast)),
    "c3t": (HESTON_EQ, 1000, {"S": 100.0, "v": 0.04}, 1 << 22, "runge-kutta", "sobol", dict(output="terminal", scramble="xor", **fast)),
    "c2": (GBM_EQ, 252, {"X1": 1.0}, 1 << 24, "euler", "sobol", dict(output="paths", scramble="xor", **fast)),
    "c2cp": (GBM_EQ, 252, {"X1": 1.0}, 1 << 24, "euler", "sobol", dict(output="paths", **fast)),
    "c2tpn": (GBM_EQ, 252, {"X1": 1.0}, 1 << 24, "euler", "sobol", dict(output="paths", layout="TPN", scramble="xor", **fast)),
    "c5": (GBM_EQ, 365, {"X1": 1.0}, 1 << 28, "euler", "pseudo", dict(output="moments", **fast)),
    "c5p": (GBM_EQ, 365, {"X1": 1.0}, 1 << 28, "euler", "pseudo", dict(output="moments", generator="philox", **fast)),
}
name = sys.argv[1]
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eqs, D, init, N, scheme, rng, kw = CFG[name]
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = int(v) if v.lstrip("-").isdigit() else v
plan = S.Plan(S.Universe(eqs, grid(D)), scheme, rng, **kw)
out = torch.empty(plan.output_shape(N), dtype=torch.float64, device="cuda")
plan.run(init, N, seed=42, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(runs):
    plan.run(init, N, seed=42, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / runs
print(f"{name} {ms:.3f} ms  {N * D / ms / 1e6:.1f} G path-steps/s  out {out.numel() * 8 / ms / 1e6:.0f} GB/s", flush=True)
