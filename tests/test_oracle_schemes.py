"""Oracle schemes/parser vs the independent Python restatement and the SURVEY.md §A.4 worked trace.

Follows src/sim/euler.rs:5-37, src/sim/runge_kutta.rs:5-107, src/func.rs:32-42, src/proc/util.rs:52-166.
"""
import math

import numpy as np
import pytest

from conftest import GBM_EQ, HESTON_EQ, grid
from oracle import py_restatement as pr


def _gbm_py():
    return [pr.Levy("X1", [(lambda c: 0.05 * c["X1"], "dt", -1, None), (lambda c: 0.1 * c["X1"], "dW", 0, None)])]


def test_worked_trace_gbm(oracle):
    # uniforms are the first three f64 of ChaCha8(seed 0) -> pseudo, seed 0, scenario 0
    U = oracle.Universe(GBM_EQ, grid(252, 3))
    e = oracle.simulate(U, {"X1": 1.0}, 1, "euler", "pseudo", seed=0)[0, :, 0]
    assert e.tolist() == [1.0, 1.0036649855005078, 1.0033247143520132, 1.0068200906983518]
    r = oracle.simulate(U, {"X1": 1.0}, 1, "runge-kutta", "pseudo", seed=0)[0, :, 0]
    assert r.tolist() == [1.0, 1.0036832451791444, 1.0033430396186047, 1.0068443835552638]
    tb = oracle.simulate(U, {"X1": 1.0}, 1, "runge-kutta", "pseudo", seed=0, textbook_rk=True)[0, :, 0]
    assert tb.tolist() == [1.0, 1.0036832451791444, 1.0033440972948497, 1.006856639483425]


@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_gbm_pseudo_matches_python(oracle, scheme):
    times = grid(252, 20)
    U = oracle.Universe(GBM_EQ, times)
    out = oracle.simulate(U, {"X1": 1.0}, 6, scheme, "pseudo", seed=42)
    for s in range(6):
        g = pr.ChaCha8(42 + s)
        raw = pr.simulate_path(_gbm_py(), times, {"X1": 1.0}, pr.StreamRng(g.next_f64, 1), scheme, 1)
        assert [r[0] for r in raw] == out[s, :, 0].tolist()


def _heston_py():
    sq = lambda c: math.pow(max(c["v"], 0.0), 0.5)
    return [
        pr.Levy("S", [(lambda c: 0.05 * c["S"], "dt", -1, None), (lambda c: sq(c) * c["S"], "dW", 0, None)]),
        pr.Levy("v", [(lambda c: 2.0 * (0.04 - c["v"]), "dt", -1, None), (lambda c: -0.21 * sq(c), "dW", 0, None),
                      (lambda c: 0.2142428528562855 * sq(c), "dW", 1, None)]),
    ]


@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
@pytest.mark.parametrize("scramble", ["cp_shift_per_path", "xor", "none"])
def test_heston_sobol_matches_python(oracle, scheme, scramble):
    times = grid(1000, 12)
    U = oracle.Universe(HESTON_EQ, times)
    assert (U.P, U.K, U.factors) == (2, 2, ["dW1", "dW2"])
    init = {"S": 100.0, "v": 0.04}
    out = oracle.simulate(U, init, 5, scheme, "sobol", seed=7, scramble=scramble)
    u = oracle.uniforms(U, 5, "sobol", seed=7, scramble=scramble)
    for s in range(5):
        raw = pr.simulate_path(_heston_py(), times, init, pr.TableRng(u[s].ravel().tolist(), 2), scheme, 2)
        assert np.allclose(np.asarray(raw), out[s], rtol=0, atol=0), (s, np.asarray(raw) - out[s])


def test_cp_shift_uniforms_follow_reference_formula(oracle):
    # sobol.rs:45-47,73-76: u = fract(point(s+5)[d] + ChaCha8(seed+s).f64[d])
    U = oracle.Universe(GBM_EQ, grid(252, 8))
    u = oracle.uniforms(U, 3, "sobol", seed=11, scramble="cp_shift_per_path")
    poly, minit = oracle.joe_kuo_from_scipy(8)
    V = pr.sobol_direction_numbers(poly, minit)
    for s in range(3):
        g = pr.ChaCha8(11 + s)
        pt = pr.sobol_point(V, s + 5)
        exp = [math.modf(x * 2.0**-64 + g.next_f64())[0] for x in pt]
        assert u[s, :, 0].tolist() == exp


def test_jump_diffusion_with_algebraic_matches_python(oracle):
    # examples/example.py:11-13 shape: state-dependent Poisson intensity + an algebraic payoff process
    eqs = ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
           "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
           "C = max(X1 - 100.0, 0.0) + X0"]
    times = grid(50, 15)
    init = {"X0": 0.3, "X1": 100.0, "C": 5.0, "unknown": 9.0}
    procs = [
        pr.Levy("X0", [(lambda c: 2.0 * (0.5 - c["X0"]), "dt", -1, None), (lambda c: 0.1, "dW", 0, None)]),
        pr.Levy("X1", [(lambda c: 0.01 * c["X1"], "dt", -1, None), (lambda c: 0.2 * c["X1"], "dW", 1, None),
                       (lambda c: 0.5 * math.cos(c["t"]), "dN", 2, lambda c: abs(c["X0"]) * 40)]),
        pr.Alg("C", lambda c: max(c["X1"] - 100.0, 0.0) + c["X0"]),
    ]
    for scheme in ("euler", "runge-kutta"):
        U = oracle.Universe(eqs, times)
        assert U.factors == ["dW1", "dW2", "dN1(abs(X0) * 40)"]
        out = oracle.simulate(U, init, 4, scheme, "pseudo", seed=3)
        for s in range(4):
            g = pr.ChaCha8(3 + s)
            raw = pr.simulate_path(procs, times, init, pr.StreamRng(g.next_f64, 3), scheme, 3)
            assert np.array_equal(np.asarray(raw), out[s]), scheme
        assert out[0, 0, 2] == 5.0                         # algebraic initial value kept at t0 (filtration.rs:42-50)


def test_algebraic_stale_cache_quirk(oracle):
    # a Levy coefficient that references an algebraic process sees its initial value at step 0
    # and 0.0 afterwards (euler.rs:31-35 + func.rs:37-39).
    eqs = ["dX = ( A ) * dt", "A = 2.0 + 0.0 * X"]
    times = [0.0, 1.0, 2.0, 3.0]
    out = oracle.simulate(oracle.Universe(eqs, times), {"X": 0.0, "A": 10.0}, 1, "euler", "pseudo", seed=0)[0]
    assert out[:, 0].tolist() == [0.0, 10.0, 10.0, 10.0]
    assert out[:, 1].tolist() == [10.0, 2.0, 2.0, 2.0]


def test_parser_acceptance_table(oracle):
    U = oracle.Universe(["delta = 1.0"], [0.0, 1.0])        # util.rs:80-82 quirk: SDE named "elta", zero terms
    assert (U.names, U.is_levy, U.num_terms) == (["elta"], [True], [0])
    U = oracle.Universe(["dX = ( 1.0 ) * dt - ( 2.0 ) * dW1"], [0.0, 1.0])   # text between terms ignored
    assert U.num_terms == [2]
    U = oracle.Universe(["dX = ( 1.0 ) * dt + ( 2.0 ) + ( 3.0 ) * dW1"], [0.0, 1.0])   # stops silently at missing '*'
    assert U.num_terms == [1] and U.K == 0
    U = oracle.Universe(["dX = ( X ) * dN1(X) + ( X ) * dN1(2*X)"], [0.0, 1.0])   # keyed by full token
    assert U.K == 2
    for bad in (["X = 1 = 2"], ["dX 1.0"], ["dX = ( 1.0 ) * dQ"], ["dX = ( 1.0 * dt"], ["dX = ( foo(1) ) * dt"],
                ["dX = ( Y ) * dt"], ["dX = ( ) * dt"]):
        with pytest.raises(ValueError):
            oracle.Universe(bad, [0.0, 1.0])


def test_expression_semantics(oracle):
    ev = oracle.expr_eval
    assert ev("2^3^2") == 512.0                              # right-assoc
    assert ev("-2^2") == 4.0                                 # unary binds to the value [3P-unverified]
    assert ev("2*3%2") == 2.0                                # % binds tighter than *
    assert ev("10/4*2") == 5.0
    assert ev("1-2-3") == -4.0 and ev("8/2/2") == 2.0
    assert ev("log(100)") == 2.0 and ev("log(2, 8)") == pytest.approx(3.0)
    assert ev("e()^1") == math.e and ev("pi()") == math.pi
    assert ev("max(1, 5, 3) + min(4, 2)") == 7.0
    assert ev("round(2.5) + int(-1.7) + sign(-3) + abs(-2)") == 3.0 - 1.0 - 1.0 + 2.0
    assert ev("1.5k + 2m") == 1500.002
    assert ev("3 > 2") == 1.0 and ev("3 < 2") == 0.0 and ev("1 and 0") == 0.0 and ev("0 or 7") == 7.0
    assert ev("x0*t", {"x0": 2.0}, t=4.0) == 8.0
    assert ev("t", {"t": 9.0}, t=4.0) == 9.0                 # a process named t shadows time (filtration.rs:72-78)
    assert ev("max(v, 0.0)^0.5 * S", {"v": 0.04, "S": 100.0}) == math.pow(0.04, 0.5) * 100.0
