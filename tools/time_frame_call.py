#!/usr/bin/env python3
"""Wall time of the drop-in call exactly as a reference script makes it (six arguments -> long DataFrame), C1 shape."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import sde_sim_rs  # noqa: E402
from conftest import GBM_EQ, grid  # noqa: E402

torch.zeros(1, device="cuda")
for i in range(4):
    t0 = time.perf_counter()
    df = sde_sim_rs.simulate(processes_equations=GBM_EQ, time_steps=grid(252), scenarios=10_000, initial_values={"X1": 1.0},
                             rng_method="pseudo", scheme="euler")
    t1 = time.perf_counter()
    r = sde_sim_rs.simulate(GBM_EQ, grid(252), 10_000, {"X1": 1.0}, "pseudo", "euler", frame=False)
    v = r.to_numpy()
    t2 = time.perf_counter()
    print(f"call {i}: frame {1e3 * (t1 - t0):.1f} ms ({type(df).__module__.split('.')[0]}, {len(df)} rows); dense to host {1e3 * (t2 - t1):.1f} ms")
