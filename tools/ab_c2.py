#!/usr/bin/env python3
"""A/B of kernel variants on the C2 workload in ONE process on one GPU (same box, same thermal state, round-robin).

usage: ab_c2.py [--rounds R] [--burst B] [--sustained S] name[:ENV=VAL[,ENV=VAL...]] ...
  each variant = environment variables read by the lowering at plan creation (SDE_B200_DEFINES="A=1;B=2", ...);
  within a variant spec use '+' instead of ';' between defines.  Extra keys: block=N (block_threads), icdf=, arith=.
Per variant and round: `burst` launches back to back after a 1.5 s pause (power controller idle), then `sustained` launches
back to back (sw_power_cap engaged); CUDA events; SM clock / power sampled over the second half of the sustained run.
The first variant's output is the reference: every other variant must reproduce it bit for bit (same arithmetic) unless
its name ends in '~' (then the max relative difference is printed instead).
"""
import argparse
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
TIMES = [k / D for k in range(D + 1)]
N = 1 << 24
ENV_KEYS = ("SDE_B200_DEFINES", "SDE_B200_RES_PIPE", "SDE_B200_NO_WIDE_TABLE")


def make_plan(spec):
    name, _, rest = spec.partition(":")
    kw = dict(icdf="fast", arith="fast", block=0)
    for k in ENV_KEYS:
        os.environ.pop(k, None)
    for item in filter(None, rest.split(",")):
        k, _, v = item.partition("=")
        if k in kw:
            kw[k] = type(kw[k])(v)
        else:
            os.environ[k] = v.replace("+", ";")
    t0 = time.perf_counter()
    plan = S.Plan(S.Universe(GBM, TIMES), "euler", "sobol", scramble="xor", icdf=kw["icdf"], arithmetic=kw["arith"], block_threads=kw["block"])
    return name, plan, time.perf_counter() - t0


def timed(plan, out, runs, sample=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(runs):
        plan.run({"X1": 1.0}, N, seed=42, out=out)
    e1.record()
    clk, pw = [], []
    if sample and nv is not None:
        while not e1.query():
            clk.append(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))
            try:
                pw.append(pynvml.nvmlDeviceGetFieldValues(nv, [pynvml.NVML_FI_DEV_POWER_INSTANT])[0].value.uiVal / 1000.0)
            except Exception:  # noqa: BLE001
                pw.append(pynvml.nvmlDeviceGetPowerUsage(nv) / 1000.0)
            time.sleep(0.004)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / runs
    half = clk[len(clk) // 2:]
    return ms, (sorted(half)[len(half) // 2] if half else None), (sum(pw[len(pw) // 2:]) / max(1, len(pw) - len(pw) // 2) if pw else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--burst", type=int, default=6)
    ap.add_argument("--sustained", type=int, default=40)
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    out = torch.empty((N, D + 1, 1), dtype=torch.float64, device="cuda")
    plans = []
    ref = None
    for spec in a.variants:
        name, plan, secs = make_plan(spec)
        plan.run({"X1": 1.0}, N, seed=42, out=out)
        torch.cuda.synchronize()
        probe = torch.cat([out[:4096].flatten(), out[N // 2:N // 2 + 4096].flatten(), out[-4096:].flatten()]).clone()
        if ref is None:
            ref, note = probe, "reference"
        elif name.endswith("~"):
            note = f"max rel diff vs first {float(((probe - ref).abs() / ref.abs()).max()):.2e}"
        else:
            note = "bit-identical to first" if torch.equal(probe, ref) else f"DIFFERS from first (max rel {float(((probe - ref).abs() / ref.abs()).max()):.2e})"
        print(f"# {name}: plan {secs:.1f} s, {note}", flush=True)
        plans.append((name, plan))
    print("variant | round | burst ms | burst G path-steps/s | sustained ms | sustained G path-steps/s | SM MHz | W", flush=True)
    for r in range(a.rounds):
        for name, plan in plans:
            time.sleep(1.5)
            plan.run({"X1": 1.0}, N, seed=42, out=out)      # one warm launch (instruction cache, tables)
            torch.cuda.synchronize()
            time.sleep(1.5)
            b_ms, _, _ = timed(plan, out, a.burst)
            s_ms, mhz, w = timed(plan, out, a.sustained, sample=True)
            print(f"{name} | {r} | {b_ms:.3f} | {N * D / b_ms / 1e6:.1f} | {s_ms:.3f} | {N * D / s_ms / 1e6:.1f} | {mhz} | {w and round(w)}", flush=True)


if __name__ == "__main__":
    main()
