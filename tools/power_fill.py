#!/usr/bin/env python3
"""Board power of plain sequential store streams (torch fill of a 34 GB f64 buffer, back to back for ~1.5 s) — the reference
point for the power the [N][T][P] scatter of the path kernel costs (tools/energy_variants.py)."""
import threading
import time

import pynvml
import torch

x = torch.empty((1 << 24) * 253, dtype=torch.float64, device="cuda")
x.fill_(1.0)
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


for label, fn in (("idle", None), ("fill f64 (sequential stores)", lambda: x.fill_(2.0))):
    samples.clear()
    stop = False
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    n = 250
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    if fn is None:
        time.sleep(1.0)
    else:
        for _ in range(n):
            fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    stop = True
    th.join()
    half = [s for s in samples if t0 + 0.5 * (t1 - t0) < s[0] < t1]
    pw = sum(s[2] for s in half) / max(1, len(half))
    clk = sorted(s[1] for s in half)[len(half) // 2] if half else 0
    if fn is None:
        print(f"{label}: {pw:.0f} W, SM {clk} MHz")
    else:
        ms = e0.elapsed_time(e1) / n
        gb = x.numel() * 8 / 1e9
        print(f"{label}: {ms:.3f} ms per {gb:.1f} GB = {gb / ms:.2f} TB/s, SM {clk} MHz, board {pw:.0f} W, {pw * ms / 1e3 / gb:.3f} J/GB")
