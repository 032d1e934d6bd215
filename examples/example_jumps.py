#!/usr/bin/env python3
"""Mean-reverting factor + jump-diffusion + an algebraic payoff process — the model family of the reference's
examples/example.py (state-dependent Poisson intensity `dN1(...)`, shared increments, algebraic rows), dense result on the GPU.

    PYTHONPATH=sde-sim-rs_b200 python examples/example_jumps.py"""
import numpy as np

import sde_sim_rs

res = sde_sim_rs.simulate(
    processes_equations=[
        "dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
        "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
        "C = max(X1 - 100.0, 0.0) + X0",
    ],
    time_steps=list(np.linspace(0.0, 1.0, 1000)),
    scenarios=10000,
    initial_values={"X0": 0.4, "X1": 100.0},
    rng_method="pseudo",
    scheme="runge-kutta",
    seed=7,
)
paths = res.values                      # torch.cuda tensor [scenarios, times, processes], reference row order
print("shape", tuple(paths.shape), "processes", res.process_names)
print("terminal means", paths[:, -1, :].mean(dim=0).tolist())
print(res.to_pandas().head(6))
