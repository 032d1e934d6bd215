"""generator="philox": the counter-based pseudo-random tier (Philox4x32-10; not in the reference).  north_star asks the
pseudo-random MC path for STATISTICAL agreement with the reference (means and variances within confidence intervals); the
stream itself is held bit-exactly to the oracle's restatement (known answers: tests/test_oracle_philox.py), so paths agree
with the oracle draw for draw as well."""
import numpy as np
import pytest
from scipy import stats

from conftest import GBM_EQ, HESTON_EQ, grid

import sde_sim_rs as S

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("eqs,init,scheme,steps", [(GBM_EQ, {"X1": 1.0}, "euler", 365), (GBM_EQ, {"X1": 1.0}, "runge-kutta", 37),
                                                  (HESTON_EQ, {"S": 100.0, "v": 0.04}, "runge-kutta", 50), (HESTON_EQ, {"S": 100.0, "v": 0.04}, "euler", 203)])
def test_philox_paths_match_the_oracle_draw_for_draw(oracle, eqs, init, scheme, steps):
    times, N, off = grid(365, steps), 1500, 2**33 + 77          # 64-bit scenario index: both counter words in use
    ref = oracle.simulate(oracle.Universe(eqs, times), init, N, scheme, "pseudo", seed=2**40 + 9, generator="philox", scenario_offset=off)
    # fast tier: 1e-11 on the GBM model (the stated end-to-end tolerance of that tier); the Heston variance process runs close to
    # zero, where the tier's absolute differences (x^0.5 by seed + Newton step, FMA contraction) are relatively larger: 1e-10
    for kw, tol in ((dict(), 1e-12), (dict(icdf="fast", arithmetic="fast"), 1e-11 if len(eqs) == 1 else 1e-10)):
        got = S.simulate(eqs, times, N, init, "pseudo", scheme, seed=2**40 + 9, generator="philox", scenario_offset=off, **kw).to_numpy()
        assert rel_err(got, ref) <= tol, (kw, rel_err(got, ref))
    term = S.simulate(eqs, times, N, init, "pseudo", scheme, seed=2**40 + 9, generator="philox", scenario_offset=off, output="terminal").to_numpy()
    assert rel_err(term, ref[:, -1, :]) <= 1e-12


def test_philox_agrees_statistically_with_the_reference_stream():
    # terminal GBM values under the two generators: same mean / variance within confidence intervals, same distribution (KS)
    times, N = grid(252), 1 << 18
    kw = dict(output="terminal", icdf="fast", arithmetic="fast")
    a = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "pseudo", "euler", seed=5, **kw).to_numpy()[:, 0]
    b = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "pseudo", "euler", seed=5, generator="philox", **kw).to_numpy()[:, 0]
    assert not np.array_equal(a, b)
    se = np.sqrt(a.var() / N + b.var() / N)
    assert abs(a.mean() - b.mean()) <= 4.5 * se                                     # means within the CI
    la, lb = np.log(a), np.log(b)
    f = la.var(ddof=1) / lb.var(ddof=1)
    lo, hi = stats.f.ppf([1e-5, 1 - 1e-5], N - 1, N - 1)
    assert lo <= f <= hi                                                            # variances within the CI
    assert stats.ks_2samp(la, lb).pvalue > 1e-4                                     # same distribution of terminal log-returns
    # against the closed form of the Euler scheme: E[X_S] = (1 + mu dt)^S
    mean = (1 + 0.05 / 252) ** 252
    assert abs(b.mean() - mean) <= 4.5 * b.std() / np.sqrt(N)
    # moments output = the same numbers, reduced on the device
    m = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "pseudo", "euler", seed=5, generator="philox", output="moments", icdf="fast", arithmetic="fast").to_numpy()[0]
    assert m[0] == N and abs(m[1] / b.mean() - 1) <= 1e-13


def test_philox_is_a_pseudo_generator_only():
    with pytest.raises(ValueError, match="unknown generator"):
        S.Plan(S.Universe(GBM_EQ, grid(252, 8)), "euler", "pseudo", generator="mt19937")
    # sobol ignores the generator (the point set is the generator)
    a = S.simulate(GBM_EQ, grid(252, 8), 64, {"X1": 1.0}, "sobol", "euler", seed=1, scramble="xor").to_numpy()
    b = S.simulate(GBM_EQ, grid(252, 8), 64, {"X1": 1.0}, "sobol", "euler", seed=1, scramble="xor", generator="philox").to_numpy()
    assert np.array_equal(a, b)
