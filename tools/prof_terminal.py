#!/usr/bin/env python3
"""One launch of each terminal-only / moments config at a reduced scenario count (for `ncu --metrics ...` captures of
the FP64 / FP32 / integer pipe utilisation the north star asks for).  usage: prof_terminal.py [scale]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import GBM_EQ, HESTON_EQ, basket_equations, grid  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
beq, binit = basket_equations(64)
CASES = [
    ("C5 GBM euler pseudo moments", GBM_EQ, grid(365), {"X1": 1.0}, (1 << 24) // scale, "euler", "pseudo", dict(output="moments")),
    ("C3 Heston RK sobol/xor terminal", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, (1 << 21) // scale, "runge-kutta", "sobol",
     dict(output="terminal", scramble="xor")),
    ("C4 basket-64 euler sobol/xor moments", beq, grid(252), binit, (1 << 17) // scale, "euler", "sobol", dict(output="moments", scramble="xor")),
]
for name, eqs, times, init, N, scheme, rng, kw in CASES:
    plan = S.Plan(S.Universe(eqs, times), scheme, rng, icdf="fast", arithmetic="fast", **kw)
    out = torch.empty(plan.output_shape(N), dtype=torch.float64, device="cuda")
    for _ in range(2):                                   # launch 0 warms up, launch 1 is the one to capture
        plan.run(init, N, seed=42, out=out)
    torch.cuda.synchronize()
    print(name, "N =", N, "steps =", len(times) - 1, flush=True)
