#!/usr/bin/env python3
"""Small runs of the tensor-core kernel (every output mode, ragged tiles) for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402
from conftest import basket_equations, grid  # noqa: E402

for n_assets, N in ((64, 300), (20, 77)):
    eqs, init = basket_equations(n_assets)
    for output in ("paths", "terminal", "moments"):
        plan = S.Plan(S.Universe(eqs, grid(252, 6)), "euler", "sobol", output=output, scramble="xor", icdf="fast", arithmetic="fast", wide_mma=2)
        out = plan.run(init, N, seed=1, scenario_offset=3)
        print(n_assets, output, tuple(out.shape), float(out.sum()), flush=True)
