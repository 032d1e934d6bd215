#!/usr/bin/env python3
"""Throughput of every BASELINE.json config on one B200 (kernel time, CUDA events, device-resident output).
The CPU restatement's rate on a bounded sample of each config comes from tests/perf_cpu_oracle.py (the oracle is test
infrastructure: only tests/, smoke() and bench.py's CPU legs execute it) and is merged in when its JSON is present.
Results -> gpurun_out/configs.json (copied to profiles/ by hand)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import GBM_EQ, HESTON_EQ, basket_equations, grid  # noqa: E402

CPU_RATES = {}
try:
    for r in json.load(open(os.path.join(ROOT, "gpurun_out", "cpu_oracle_rates.json"))):
        CPU_RATES[r["config"]] = r
except Exception:  # noqa: BLE001
    pass


def gpu_time(plan, init, N, reps=3, **kw):
    out = torch.empty(plan.output_shape(N), dtype=torch.float64, device="cuda")
    for _ in range(2):
        plan.run(init, N, seed=42, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run(init, N, seed=42, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    del out
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) / reps


rows = []


def run(name, eqs, times, init, N, scheme, rng, n_cpu, plan_kw, cpu_kw=None, bytes_per_path_step=None):
    S_ = len(times) - 1
    for label, extra in (("reference-arith", dict(icdf="reference", arithmetic="strict")), ("fast", dict(icdf="fast", arithmetic="fast"))):
        plan = S.Plan(S.Universe(eqs, times), scheme, rng, **plan_kw, **extra)
        ms = gpu_time(plan, init, N)
        r = {"config": name, "mode": label, "N": N, "steps": S_, "scheme": scheme, "rng": rng, **{k: v for k, v in plan_kw.items()},
             "ms": round(ms, 3), "path_steps_per_s": N * S_ / ms * 1e3}
        if bytes_per_path_step:
            r["out_GBps"] = round(N * S_ * bytes_per_path_step / ms / 1e6, 1)
        rows.append(r)
        print(json.dumps(r), flush=True)
    if name in CPU_RATES:
        rows.append(CPU_RATES[name])
        print(json.dumps(CPU_RATES[name]), flush=True)


run("C1 GBM euler pseudo 10k x 252 full paths", GBM_EQ, grid(252), {"X1": 1.0}, 10_000, "euler", "pseudo", 10_000,
    dict(output="paths"), bytes_per_path_step=8)
run("C2 GBM euler sobol/xor 2^24 x 252 full paths", GBM_EQ, grid(252), {"X1": 1.0}, 1 << 24, "euler", "sobol", 1 << 19,
    dict(output="paths", scramble="xor"), dict(scramble="xor"), bytes_per_path_step=8)
run("C2' same, reference cp_shift_per_path scramble", GBM_EQ, grid(252), {"X1": 1.0}, 1 << 24, "euler", "sobol", 1 << 18,
    dict(output="paths", scramble="cp_shift_per_path"), dict(scramble="cp_shift_per_path"), bytes_per_path_step=8)
run("C3 Heston RK sobol/xor 2^22 x 1000 full paths", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 1 << 22, "runge-kutta", "sobol", 1 << 16,
    dict(output="paths", scramble="xor"), dict(scramble="xor"), bytes_per_path_step=16)
run("C3 Heston RK sobol/xor 2^22 x 1000 terminal", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 1 << 22, "runge-kutta", "sobol", 1 << 16,
    dict(output="terminal", scramble="xor"), dict(scramble="xor"))
run("C3 Heston RK pseudo 2^22 x 1000 terminal", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 1 << 22, "runge-kutta", "pseudo", 1 << 16,
    dict(output="terminal"))
eqs, init = basket_equations(64)
run("C4 basket-64 euler sobol/xor 2^20 x 252 moments", eqs, grid(252), init, 1 << 20, "euler", "sobol", 1 << 11,
    dict(output="moments", scramble="xor"), dict(scramble="xor"))
run("C5 GBM euler pseudo 2^30 x 365 moments (1/8 of 2^33)", GBM_EQ, grid(365), {"X1": 1.0}, 1 << 30, "euler", "pseudo", 1 << 19,
    dict(output="moments"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
