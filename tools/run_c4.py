#!/usr/bin/env python3
"""Time the C4 workload (64-asset correlated GBM basket, RQMC Sobol, 2^20 paths x 252 steps, moments).
usage: run_c4.py [N] [wide_mma,block[,min_blocks] ...]   wide_mma: 0 auto (tensor-core kernel), 1 time-tiled kernel"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import basket_equations, grid  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
specs = sys.argv[2:] or ["0,0", "1,0"]
RUNS = int(os.environ.get("RUNS", 2))
OUTPUT = os.environ.get("OUTPUT", "moments")
eqs, init = basket_equations(int(os.environ.get("ASSETS", 64)))
for spec in specs:
    f = [int(x) for x in spec.split(",")] + [0, 0]
    try:
        plan = S.Plan(S.Universe(eqs, grid(252)), "euler", "sobol", output=OUTPUT, scramble="xor", icdf=os.environ.get("ICDF", "fast"),
                      arithmetic="fast", wide_mma=f[0], block_threads=f[1], min_blocks=f[2])
        kind = "wide(dmma)" if "sde_sim_wide.cuh" in plan.source else "tiled"
        out = torch.empty(plan.output_shape(N), dtype=torch.float64, device="cuda")
        plan.run(init, N, seed=42, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(RUNS):
            plan.run(init, N, seed=42, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / RUNS
        print(spec, kind, f"{ms:.1f} ms  {N * 252 / ms / 1e6:.3f} G path-steps/s  {N * 252 * len(init) / ms / 1e6:.1f} G asset-steps/s", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(spec, "error", str(ex)[:300], flush=True)
