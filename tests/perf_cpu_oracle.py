#!/usr/bin/env python3
"""Rate of the CPU oracle (C++ restatement of the reference algorithm, OpenMP over scenarios like rayon in
src/sim/mod.rs:41-43) on a bounded sample of every BASELINE.json config, all host cores.  Not a test (not collected):
a measurement helper that lives under tests/ because only tests/, smoke() and bench.py may execute oracle/.
Writes gpurun_out/cpu_oracle_rates.json, which tools/bench_configs.py merges into its table."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import GBM_EQ, HESTON_EQ, basket_equations, grid  # noqa: E402
from oracle import oracle as orc  # noqa: E402

THREADS = len(os.sched_getaffinity(0))
beq, binit = basket_equations(64)
CASES = [
    ("C1 GBM euler pseudo 10k x 252 full paths", GBM_EQ, grid(252), {"X1": 1.0}, 10_000, "euler", "pseudo", {}),
    ("C2 GBM euler sobol/xor 2^24 x 252 full paths", GBM_EQ, grid(252), {"X1": 1.0}, 1 << 19, "euler", "sobol", {"scramble": "xor"}),
    ("C2' same, reference cp_shift_per_path scramble", GBM_EQ, grid(252), {"X1": 1.0}, 1 << 18, "euler", "sobol", {"scramble": "cp_shift_per_path"}),
    ("C3 Heston RK sobol/xor 2^22 x 1000 full paths", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 1 << 16, "runge-kutta", "sobol", {"scramble": "xor"}),
    ("C3 Heston RK sobol/xor 2^22 x 1000 terminal", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 1 << 16, "runge-kutta", "sobol", {"scramble": "xor"}),
    ("C3 Heston RK pseudo 2^22 x 1000 terminal", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 1 << 16, "runge-kutta", "pseudo", {}),
    ("C4 basket-64 euler sobol/xor 2^20 x 252 moments", beq, grid(252), binit, 1 << 11, "euler", "sobol", {"scramble": "xor"}),
    ("C5 GBM euler pseudo 2^30 x 365 moments (1/8 of 2^33)", GBM_EQ, grid(365), {"X1": 1.0}, 1 << 19, "euler", "pseudo", {}),
]
if __name__ == "__main__":
    orc.build()
    rows = []
    for name, eqs, times, init, N, scheme, rng, kw in CASES:
        U = orc.Universe(eqs, times)
        t0 = time.perf_counter()
        orc.simulate(U, init, N, scheme, rng, seed=42, nthreads=THREADS, **kw)
        rate = N * (len(times) - 1) / (time.perf_counter() - t0)
        rows.append({"config": name, "mode": "cpu-oracle", "N": N, "steps": len(times) - 1, "threads": THREADS, "path_steps_per_s": rate})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "cpu_oracle_rates.json"), "w"), indent=1)
