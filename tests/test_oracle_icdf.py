"""Known answers for the in-tree inverse CDFs (src/proc/increment.rs:160-200)."""
import numpy as np
from scipy.stats import norm


def test_icdf_normal_kats(oracle):
    # SURVEY.md §B.4 (glibc log/sqrt)
    kats = {0.5: -1.0100667546808495e-07, 0.975: 1.9603949169253396, 0.025: -1.96039491692534,
            0.875: 1.150435626267757, 0.375: -0.3181998624529798, 1e-9: -5.997437910337459,
            0.7090754154265618: 0.5503013721626198}
    got = oracle.icdf_normal(list(kats))
    assert got.tolist() == list(kats.values())


def test_icdf_normal_error_vs_true_quantile(oracle):
    p = np.linspace(1e-4, 1 - 1e-4, 10001)
    err = np.abs(oracle.icdf_normal(p) - norm.ppf(p))
    assert err.max() < 4.5e-4           # Abramowitz–Stegun 26.2.23 bound


def test_icdf_normal_p0_is_nan(oracle):
    assert np.isnan(oracle.icdf_normal([0.0])[0])   # ln 0 = -inf -> inf - inf/inf (increment.rs:165-177)


def test_icdf_poisson_kats(oracle):
    assert [oracle.icdf_poisson(u, 0.05) for u in (0.1, 0.95, 0.96, 0.999, 0.9999)] == [0, 0, 1, 2, 2]
    assert [oracle.icdf_poisson(u, 3.0) for u in (0.01, 0.5, 0.99)] == [0, 3, 8]
    assert oracle.icdf_poisson(0.5, 0.0) == 0 and oracle.icdf_poisson(0.5, -1.0) == 0
    assert oracle.icdf_poisson(1.0, 500.0) == 200        # cap (increment.rs:193)


def test_python_restatement_agrees(oracle):
    from oracle import py_restatement as pr

    p = np.random.default_rng(0).random(2000)
    assert oracle.icdf_normal(p).tolist() == [pr.icdf_normal(float(x)) for x in p]
