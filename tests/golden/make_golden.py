#!/usr/bin/env python3
"""Generates tests/golden/golden_v1.npz — the committed fixtures the parity tests compare against.

The reference (a Rust crate) cannot be built or run in this image and ships no tests or golden vectors of its own
(SURVEY.md §4, §8c), so nothing here comes from the reference binary.  Two kinds of fixture instead:

  * `pin_*`   : values from sources INDEPENDENT of this repo's oracle that its third-party arithmetic must reproduce —
                scipy's Sobol engine (same Joe–Kuo table and Gray-code order as the `sobol` crate), the ChaCha8 zero-key
                known answer and rand_chacha's construction test, A&S 26.2.23 evaluated with Python's `math` (glibc), the
                SURVEY.md §A.4 worked Runge–Kutta trace.
  * `case_*`  : outputs of the CPU oracle (oracle/sde_oracle.cpp) for every BASELINE.json config shape at a small scenario
                count, frozen so that (a) the oracle cannot drift silently and (b) the `-m gpu` tests can compare the CUDA
                path with committed numbers.  Their inputs are listed in CASES below (seeded, deterministic).

    python tests/golden/make_golden.py          # rewrites golden_v1.npz (review the diff before committing)"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import GBM_EQ, HESTON_EQ, basket_equations, grid  # noqa: E402

JUMP_EQ = ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
           "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
           "C = max(X1 - 100.0, 0.0) + X0"]
_beq, _binit = basket_equations(64)
# name -> (equations, times, initial values, scenarios, scheme, rng_method, oracle keyword arguments)
CASES = {
    "c1_gbm_euler_pseudo": (GBM_EQ, grid(252), {"X1": 1.0}, 8, "euler", "pseudo", dict(seed=1)),
    "c2_gbm_euler_sobol_xor": (GBM_EQ, grid(252), {"X1": 1.0}, 8, "euler", "sobol", dict(seed=42, scramble="xor")),
    "c2_gbm_euler_sobol_cp_shift": (GBM_EQ, grid(252), {"X1": 1.0}, 8, "euler", "sobol", dict(seed=42, scramble="cp_shift_per_path")),
    "c2_gbm_euler_sobol_offset": (GBM_EQ, grid(252, 40), {"X1": 1.0}, 6, "euler", "sobol", dict(seed=9, scramble="xor", scenario_offset=1_000_003)),
    "c3_heston_rk_sobol_xor": (HESTON_EQ, grid(1000, 60), {"S": 100.0, "v": 0.04}, 6, "runge-kutta", "sobol", dict(seed=7, scramble="xor")),
    "c3_heston_rk_pseudo": (HESTON_EQ, grid(1000, 60), {"S": 100.0, "v": 0.04}, 6, "runge-kutta", "pseudo", dict(seed=7)),
    "c4_basket64_euler_sobol_xor": (_beq, grid(252, 12), _binit, 5, "euler", "sobol", dict(seed=3, scramble="xor")),
    "c5_gbm_euler_pseudo_365": (GBM_EQ, grid(365), {"X1": 1.0}, 8, "euler", "pseudo", dict(seed=77)),
    "jump_rk_pseudo": (JUMP_EQ, grid(50, 15), {"X0": 0.4, "X1": 100.0}, 6, "runge-kutta", "pseudo", dict(seed=5)),
}


def as_icdf(p):
    """Abramowitz–Stegun 26.2.23 with the reference's constants, written out again from the formula
    (src/proc/increment.rs:160-179) and evaluated by Python's math (glibc log / sqrt)."""
    c0, c1, c2, d1, d2, d3 = 2.515517, 0.802853, 0.010328, 1.432788, 0.189269, 0.001308
    w = p if p < 0.5 else 1.0 - p
    t = math.sqrt(-2.0 * math.log(w))
    x = t - ((c2 * t + c1) * t + c0) / (((d3 * t + d2) * t + d1) * t + 1.0)
    return -x if p < 0.5 else x


def main():
    from scipy.stats import qmc

    out = {}
    # ---- independent pins
    out["pin_sobol_scipy_d252_first64"] = (qmc.Sobol(d=252, scramble=False, bits=64).random(64) * 2.0**64).astype(np.uint64)
    eng = qmc.Sobol(d=16, scramble=False, bits=32)
    eng.fast_forward((1 << 20) + 12345)
    out["pin_sobol_scipy_d16_from_1060921"] = (eng.random(8) * 2.0**64).astype(np.uint64)
    out["pin_chacha8_zero_key_block0"] = np.frombuffer(bytes.fromhex(
        "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
        "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42"), dtype=np.uint8).copy()
    out["pin_rand_chacha_construction_first_u32"] = np.array([137206642], dtype=np.uint32)
    ps = np.concatenate([2.0 ** -np.arange(1, 54, 4.0), 1.0 - 2.0 ** -np.arange(2, 54, 4.0), np.linspace(0.001, 0.999, 41)])
    out["pin_icdf_p"] = ps
    out["pin_icdf_as_glibc"] = np.array([as_icdf(float(p)) for p in ps])
    # ---- frozen oracle outputs
    from oracle import oracle as orc

    orc.build()
    for name, (eqs, times, init, n, scheme, rng, kw) in CASES.items():
        out["case_" + name] = orc.simulate(orc.Universe(eqs, times), init, n, scheme, rng, **kw)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    for k, v in out.items():
        print(f"{k:45s} {str(v.dtype):8s} {v.shape}")


if __name__ == "__main__":
    main()
