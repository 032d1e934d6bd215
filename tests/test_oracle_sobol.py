"""Pins the oracle's Sobol restatement (sobol crate 1.0.2, src/rng/sobol.rs:15-25) to scipy."""
import numpy as np
import pytest
from scipy.stats import qmc


def test_direct_equals_sequential(oracle):
    V = oracle.sobol_direction_numbers(40)
    a = oracle.sobol_points(V, 0, 300, sequential=True)
    b = oracle.sobol_points(V, 0, 300, sequential=False)
    assert np.array_equal(a, b)
    a = oracle.sobol_points(V, 1000, 64, sequential=True)
    b = oracle.sobol_points(V, 1000, 64, sequential=False)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dims", [4, 252, 2000])
def test_points_match_scipy_bits64(oracle, dims):
    V = oracle.sobol_direction_numbers(dims)
    n = 64
    pts = oracle.sobol_points(V, 0, n)
    ref = qmc.Sobol(d=dims, scramble=False, bits=64).random(n)       # Gray-code order from the origin
    mine = pts.astype(np.float64) * 2.0**-64                          # exact: only the top 6 bits are set
    assert np.array_equal(mine, ref)


def test_survey_known_points(oracle):
    # SURVEY.md §B.1 [3P-verified-here] values; index 5 is the first point the reference uses (skip(5)).
    V = oracle.sobol_direction_numbers(16128)
    p = oracle.sobol_points(V, 0, 8).astype(np.float64) * 2.0**-64
    assert p[5, :4].tolist() == [0.875, 0.875, 0.125, 0.375]
    assert p[4, :4].tolist() == [0.375, 0.375, 0.625, 0.875]
    assert p[5, 249:252].tolist() == [0.875, 0.625, 0.625]
    assert p[5, 1997:2000].tolist() == [0.125, 0.875, 0.875]
    assert p[5, 16125:16128].tolist() == [0.375, 0.375, 0.875]


def test_high_index_matches_scipy_fast_forward(oracle):
    dims = 16
    V = oracle.sobol_direction_numbers(dims)
    start = (1 << 20) + 12345
    # scipy's fast_forward rejects bits=64 (dtype bug); for n < 2^32 only the top 32 bits of the
    # 64-bit integers are set, so the bits=32 engine renders the same f64 values.
    eng = qmc.Sobol(d=dims, scramble=False, bits=32)
    eng.fast_forward(start)
    ref = eng.random(8)
    mine = oracle.sobol_points(V, start, 8).astype(np.float64) * 2.0**-64
    assert np.array_equal(mine, ref)


def test_python_restatement_agrees(oracle):
    from oracle import py_restatement as pr

    poly, minit = oracle.joe_kuo_from_scipy(12)
    Vp = pr.sobol_direction_numbers(poly, minit)
    V = oracle.sobol_direction_numbers(12)
    assert [[int(x) for x in row] for row in V] == Vp
    for n in (0, 1, 5, 6, 1023, 99999):
        assert pr.sobol_point(Vp, n) == [int(x) for x in oracle.sobol_points(V, n, 1)[0]]
