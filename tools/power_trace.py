#!/usr/bin/env python3
"""Clock / power trace of the C2 kernel under back-to-back launches (evidence for DESIGN.md §4.1d "board power"):
per-launch kernel times from CUDA events next to NVML samples of SM clock, instantaneous board power and throttle reasons.
usage: power_trace.py [launches] [out.csv]    (default 60 launches -> gpurun_out/power_trace.csv)"""
import os
import sys
import threading
import time

import pynvml
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sde-sim-rs_b200"))
import sde_sim_rs as S  # noqa: E402

launches = int(sys.argv[1]) if len(sys.argv) > 1 else 60
out_csv = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "power_trace.csv")
GBM = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
D, N = 252, 1 << 24
plan = S.Plan(S.Universe(GBM, [k / D for k in range(D + 1)]), "euler", "sobol", scramble="xor", icdf="fast", arithmetic="fast")
out = torch.empty((N, D + 1, 1), dtype=torch.float64, device="cuda")
plan.run({"X1": 1.0}, N, seed=42, out=out)
torch.cuda.synchronize()
time.sleep(1.0)                                              # start from an idle power controller

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
samples, stop = [], False


def sampler():
    while not stop:
        t = time.perf_counter()
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        try:
            pw = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_POWER_INSTANT])[0].value.uiVal / 1000.0
        except Exception:  # noqa: BLE001
            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
            else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        samples.append((t, clk, pw, int(reasons)))
        time.sleep(0.002)


th = threading.Thread(target=sampler, daemon=True)
th.start()
time.sleep(0.05)
events = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
t_start = time.perf_counter()
events[0].record()
for i in range(launches):
    plan.run({"X1": 1.0}, N, seed=42, out=out)
    events[i + 1].record()
torch.cuda.synchronize()
t_end = time.perf_counter()
time.sleep(0.05)
stop = True
th.join()
ms = [events[i].elapsed_time(events[i + 1]) for i in range(launches)]
SW_POWER_CAP = 0x4
os.makedirs(os.path.dirname(out_csv), exist_ok=True)
with open(out_csv, "w") as f:
    f.write("# C2 kernel, %d launches back to back; kernel_ms per launch (CUDA events); NVML samples every ~2 ms\n" % launches)
    f.write("launch,kernel_ms,G_path_steps_per_s\n")
    for i, m in enumerate(ms):
        f.write("%d,%.4f,%.1f\n" % (i, m, N * D / m / 1e6))
    f.write("t_ms,sm_mhz,power_w,sw_power_cap\n")
    for t, clk, pw, r in samples:
        if t_start - 0.02 <= t <= t_end + 0.02:
            f.write("%.1f,%d,%.0f,%d\n" % ((t - t_start) * 1e3, clk, pw, 1 if r & SW_POWER_CAP else 0))
first, last = ms[:5], ms[-10:]
inr = [(t, c, p, r) for t, c, p, r in samples if t_start <= t <= t_end]
cap_t = next(((t - t_start) * 1e3 for t, c, p, r in inr if r & SW_POWER_CAP), None)
print("first 5 launches: %.3f ms (%.1f G)   last 10: %.3f ms (%.1f G)" % (sum(first) / 5, N * D / (sum(first) / 5) / 1e6, sum(last) / 10, N * D / (sum(last) / 10) / 1e6))
print("sw_power_cap first seen at %s ms; SM clock min %d MHz, max power %.0f W, %d samples" % (
    "%.0f" % cap_t if cap_t is not None else "never", min(c for _, c, _, _ in inr), max(p for _, _, p, _ in inr), len(inr)))
