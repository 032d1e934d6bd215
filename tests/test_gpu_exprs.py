"""Every operator and builtin of the fasteval subset through the device lowering (csrc/host/expr.cpp ->
sde_expr_helpers.cuh), evaluated on the GPU as algebraic processes and as jump intensities dN(lambda(t, X)), against the CPU
oracle's evaluator (orc_expr_eval, the restatement of what src/func.rs:18-42 hands to fasteval).

Operators that contain '=' (==, !=, >=, <=) cannot reach the evaluator: parse_single_equation splits the whole equation
string on '=' and rejects more than two parts (src/proc/util.rs:73-76) — asserted below.

Bar: strict arithmetic, <= 4 ulp (CUDA's sin / cos / log / pow vs glibc's can differ by an ulp or two; + - * / % are
correctly rounded on both sides and must agree exactly)."""
import numpy as np
import pytest

import sde_sim_rs as S

pytestmark = pytest.mark.gpu

VARS = {"x": 1.75, "y": -0.6, "z": 3.0}
EXPRS = [
    # arithmetic, precedence (each binary operator its own level, ^ right-associative), unary operators
    "1 + 2 * 3", "2 ^ 3 ^ 2", "-2 ^ 2", "2 * -x", "x - y - z", "x / y / z", "x % y", "7.5 % 2", "-7.5 % 2", "x ^ y", "x ^ 0.5", "x ^ 2", "x ^ -1",
    "1 - 2 - 3 * 4 / 5", "(x + y) * (x - y)", "x * y + z", "x + y * z", "x * x * x", "1 / 3", "x / 3 * 3", "+x", "-(x + y)", "(0 - 1) ^ 0.5",
    # comparisons and logic (1.0 / 0.0; `and` / `or` return an operand)
    "x < y", "x > y", "y < x", "x < y or z > 2", "x > y and z < 2", "!(x > y)", "!0", "!2", "x > 1 and y", "0 or y", "x < y || z > 2", "x > y && z > 2",
    # literals with SI-style suffixes and exponents
    "1k", "2.5M", "3m", "4u", "5n", "6p", "1.5e3", "1e-3 * x", "2G / 1T",
    # builtins
    "abs(y)", "sign(y)", "sign(x)", "int(x)", "int(y)", "int(-1.5)", "ceil(x)", "ceil(y)", "floor(x)", "floor(y)",
    "round(x)", "round(y)", "round(1.5)", "round(0.5, x)", "round(0.1, 0.26)",
    "log(x)", "log(100)", "log(2, 8)", "log(e(), x)", "min(x, y)", "max(x, y)", "min(x, y, z)", "max(x, y, z)", "max(x - 100.0, 0.0)",
    "e()", "pi()", "sin(x)", "cos(x)", "tan(x)", "asin(y)", "acos(y)", "atan(x)", "sinh(x)", "cosh(x)", "tanh(x)", "asinh(x)", "acosh(x)", "atanh(y)",
    # the expressions of the reference's own examples (examples/example.rs:12-13, example.py:11-13, example_gbm.py:14) and time
    "sin(t)", "0.5 * cos(t)", "0.01 * x", "2.0 * (0.5 - x)", "0.05 * x", "0.1 * x", "e() ^ x", "max(y, 0.0) ^ 0.5 * x", "-0.21 * max(z, 0.0) ^ 0.5", "t", "t + x",
]


def _ulps(a, b):
    if np.isnan(a) and np.isnan(b):
        return 0.0
    if a == b:
        return 0.0
    return abs(a - b) / np.spacing(abs(b)) if np.isfinite(a) and np.isfinite(b) else np.inf


@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_builtins_and_operators_as_algebraic_processes(oracle, scheme):
    eqs = ["dx = ( 0.0 ) * dt + ( 0.0 ) * dW1", "dy = ( 0.0 ) * dt", "dz = ( 0.0 ) * dt"] + [f"A{i} = {e}" for i, e in enumerate(EXPRS)]
    times = [0.0, 0.25, 1.0]
    got = S.simulate(eqs, times, 3, VARS, "pseudo", scheme, seed=1).to_numpy()
    assert got.shape == (3, 3, 3 + len(EXPRS))
    worst, exact_ops = 0.0, ("+", "-", "*", "/", "%")
    for i, e in enumerate(EXPRS):
        for ti in (1, 2):                                     # algebraic processes are evaluated from step 1 on (euler.rs:31-36)
            want = oracle.expr_eval(e, VARS, times[ti])
            have = float(got[0, ti, 3 + i])
            d = _ulps(have, want)
            worst = max(worst, d)
            assert d <= 4, (e, ti, have, want, d)
            assert got[1, ti, 3 + i] == have or (np.isnan(have) and np.isnan(got[1, ti, 3 + i]))
    print("expressions:", len(EXPRS), "worst ulp distance", worst)
    # whole-model parity with the oracle on top (the same expressions through its simulate)
    ref = oracle.simulate(oracle.Universe(eqs, times), VARS, 3, scheme, "pseudo", seed=1)
    both_nan = np.isnan(ref) & np.isnan(got)
    assert np.all(both_nan | (np.abs(got - ref) <= 4 * np.spacing(np.abs(ref))))


def test_expressions_as_jump_intensities(oracle):
    lam = ["abs(y) * 40", "z ^ 2", "max(x, z) * 10", "log(2, 8) * 5", "30 * (x > y)", "sin(t) ^ 2 * 50 + 1", "min(x, y, z) + 5", "2k / 100", "round(0.5, x) * 7", "e() ^ x"]
    eqs = ["dx = ( 0.0 ) * dt", "dy = ( 0.0 ) * dt", "dz = ( 0.0 ) * dt"] + [f"dJ{i} = ( 1.0 ) * dN{i}({e})" for i, e in enumerate(lam)]
    times = [k / 20 for k in range(21)]
    init = dict(VARS, **{f"J{i}": 0.0 for i in range(len(lam))})
    N = 400
    ref = oracle.simulate(oracle.Universe(eqs, times), init, N, "euler", "pseudo", seed=5)
    got = S.simulate(eqs, times, N, init, "pseudo", "euler", seed=5).to_numpy()
    assert ref[:, -1, 3:].sum() > 1000                        # the intensities are live: thousands of jumps
    # a Poisson count can flip where u sits within an ulp of a CDF step (CUDA exp vs glibc exp): compare per path
    bad = np.flatnonzero(np.any(got != ref, axis=(1, 2)))
    assert bad.size <= 2, bad.size


def test_operators_with_equals_sign_never_reach_the_evaluator():
    for op in ("==", "!=", ">=", "<="):
        with pytest.raises(ValueError, match="Failed to parse equations"):
            S.Universe([f"A = x {op} 1"], [0.0, 1.0])
