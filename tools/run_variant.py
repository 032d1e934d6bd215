#!/usr/bin/env python3
"""Run the C2 workload once per launch shape given on the command line (for ncu captures and quick timings).
usage: run_variant.py N direct,block,min_blocks[,tile] [direct,block,min_blocks ...]   env RUNS=launches per variant"""
import os
import sys

import time

import torch

try:
    import pynvml
    pynvml.nvmlInit()
    nv = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:  # noqa: BLE001
    pynvml, nv = None, None

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sde-sim-rs_b200"))
import sde_sim_rs as S  # noqa: E402

GBM = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
D = 252
times = [k / D for k in range(D + 1)]
N = int(sys.argv[1])
runs = int(os.environ.get("RUNS", 3))
DTYPE = os.environ.get("DTYPE", "f64")
out = torch.empty((N, D + 1, 1), dtype=torch.float32 if DTYPE == "f32" else torch.float64, device="cuda")
for spec in sys.argv[2:]:
    f = [int(x) for x in spec.split(",")] + [0]
    plan = S.Plan(S.Universe(GBM, times), "euler", "sobol", scramble="xor", ntp_direct=f[0], block_threads=f[1], min_blocks=f[2],
                  tile_steps=f[3], icdf=os.environ.get("ICDF", "fast"), arithmetic="fast", dtype=DTYPE)
    plan.run({"X1": 1.0}, N, seed=42, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(runs):
        plan.run({"X1": 1.0}, N, seed=42, out=out)
    e1.record()
    clk, pw = [], []
    if nv is not None:                                   # launches are asynchronous: sample clocks / power while they run
        while not e1.query():
            clk.append(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))
            try:
                pw.append(pynvml.nvmlDeviceGetFieldValues(nv, [pynvml.NVML_FI_DEV_POWER_INSTANT])[0].value.uiVal / 1000.0)
            except Exception:  # noqa: BLE001
                pw.append(pynvml.nvmlDeviceGetPowerUsage(nv) / 1000.0)
            time.sleep(0.05)                             # (polling NVML at 200 Hz slowed the launches by 25 %: 20 Hz)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / runs
    extra = ""
    if clk:
        half = clk[len(clk) // 2:]                       # second half of the run: the power controller has settled
        extra = f"  sm_mhz median(2nd half) {sorted(half)[len(half) // 2]}  power W mean(2nd half) {sum(pw[len(pw) // 2:]) / len(half):.0f}"
    print(spec, f"{ms:.3f} ms  {N * D / ms / 1e6:.1f} G path-steps/s{extra}", flush=True)
