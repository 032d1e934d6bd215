#!/usr/bin/env python3
"""Time the C4 workload (64-asset correlated GBM basket, RQMC Sobol, 2^20 paths x 252 steps, moments)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import basket_equations, grid  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
eqs, init = basket_equations(64)
for kw in (dict(icdf="fast", arithmetic="fast"),):
    shapes = ((0, 0), (256, 2), (256, 3), (128, 4), (128, 6), (512, 1), (512, 2)) if not os.environ.get("C4_ONE") else ((0, 0),)
    for block, mb in shapes:
        try:
            plan = S.Plan(S.Universe(eqs, grid(252)), "euler", "sobol", output="moments", scramble="xor", block_threads=block, min_blocks=mb, **kw)
            out = torch.empty(plan.output_shape(N), dtype=torch.float64, device="cuda")
            plan.run(init, N, seed=42, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                plan.run(init, N, seed=42, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            print(block, mb, f"{ms:.1f} ms  {N * 252 / ms / 1e6:.3f} G path-steps/s  {N * 252 * 64 / ms / 1e6:.1f} G asset-steps/s", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(block, mb, "error", str(ex)[:200])
