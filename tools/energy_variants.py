#!/usr/bin/env python3
"""Sustained clock / board power / time per launch of one config (back-to-back launches, NVML sampled at 20 Hz so that the
polling does not disturb the launches).  Run once per variant — the SDE_B200_DEBUG_* / SDE_B200_DEFINES environment is read at
plan creation:   energy_variants.py <config of run_cfg.py> [launches]"""
import os
import subprocess
import sys
import threading
import time

import pynvml
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import GBM_EQ, HESTON_EQ, grid  # noqa: E402

fast = dict(icdf="fast", arithmetic="fast")
CFG = {
    "c2": (GBM_EQ, 252, {"X1": 1.0}, 1 << 24, "euler", "sobol", dict(output="paths", scramble="xor", **fast)),
    "c3": (HESTON_EQ, 1000, {"S": 100.0, "v": 0.04}, 1 << 22, "runge-kutta", "sobol", dict(output="paths", scramble="xor", **fast)),
}
name = sys.argv[1]
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 150
eqs, D, init, N, scheme, rng, kw = CFG[name]
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = int(v) if v.lstrip("-").isdigit() else v
plan = S.Plan(S.Universe(eqs, grid(D)), scheme, rng, **kw)
out = torch.empty(plan.output_shape(N), dtype=torch.float64, device="cuda")
plan.run(init, N, seed=42, out=out)
torch.cuda.synchronize()
time.sleep(1.0)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        try:
            pw = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_POWER_INSTANT])[0].value.uiVal / 1000.0
        except Exception:  # noqa: BLE001
            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        samples.append((time.perf_counter(), clk, pw))
        time.sleep(0.05)


th = threading.Thread(target=sampler, daemon=True)
th.start()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
t0 = time.perf_counter()
ev[0].record()
for i in range(launches):
    plan.run(init, N, seed=42, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
stop = True
th.join()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(launches)]
half = [s for s in samples if s[0] > t0 + 0.5 * (t1 - t0) and s[0] < t1]
tail = ms[launches // 2:]
clk = sorted(s[1] for s in half)[len(half) // 2] if half else 0
pw = sum(s[2] for s in half) / max(1, len(half))
tag = " ".join([f"{k}={v}" for k, v in os.environ.items() if k.startswith("SDE_B200_")] + sys.argv[3:])
print(f"{name} [{tag or 'default'}]: first 5 launches {sum(ms[:5]) / 5:.3f} ms; second half {sum(tail) / len(tail):.3f} ms/launch, "
      f"SM {clk} MHz, board {pw:.0f} W, {pw * sum(tail) / len(tail) / 1e3:.2f} J/launch ({len(half)} NVML samples)", flush=True)
