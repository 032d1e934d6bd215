#!/usr/bin/env python3
"""Tuning sweep on a GPU box: time the fused kernel on the C2 workload (reduced N) over launch-shape options
and print path-steps/s; used to pick the lowering defaults."""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sde-sim-rs_b200"))
import sde_sim_rs as S  # noqa: E402

GBM = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
D = 252
times = [k / D for k in range(D + 1)]
N = int(os.environ.get("SWEEP_N", 1 << 22))
modes = {"fast": dict(icdf="fast", arithmetic="fast"), "strict": dict(icdf="reference", arithmetic="strict")}
which = sys.argv[1:] or ["fast"]
out = torch.empty((N, D + 1, 1), dtype=torch.float64, device="cuda")
res = []
grid = [("NTP", 2, 0, 256, 2), ("NTP", 2, 0, 256, 4)]                     # time-tiled kernel, direct sector stores
grid += [("NTP", 3, 0, b, 1) for b in (128, 256, 384, 512, 768, 1024)]   # persistent kernel, one CTA of b threads per SM
if os.environ.get("SWEEP_GRID"):
    grid = [tuple(("NTP",) + tuple(int(x) for x in g.split(","))) for g in os.environ["SWEEP_GRID"].split(";")]
for mode in which:
    for layout, direct, tt, block, mb in grid:
        try:
            plan = S.Plan(S.Universe(GBM, times), "euler", "sobol", scramble="xor", layout=layout, tile_steps=tt,
                          block_threads=block, min_blocks=mb, ntp_direct=direct, **modes[mode])
            o = out if layout == "NTP" else out.view(D + 1, 1, N)
            for _ in range(2):
                plan.run({"X1": 1.0}, N, seed=42, out=o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = int(os.environ.get("SWEEP_REPS", 40))    # long enough for the power controller to settle
            for _ in range(reps):
                plan.run({"X1": 1.0}, N, seed=42, out=o)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            r = {"mode": mode, "layout": layout, "direct": direct, "tt": tt, "block": block, "min_blocks": mb, "ms": round(ms, 4),
                 "gps": round(N * D / ms / 1e6, 1), "gbs": round(N * (D + 1) * 8 / ms / 1e6, 1)}
        except Exception as ex:  # noqa: BLE001
            r = {"mode": mode, "layout": layout, "direct": direct, "tt": tt, "block": block, "min_blocks": mb, "error": str(ex)[:200]}
        print(json.dumps(r), flush=True)
        res.append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
