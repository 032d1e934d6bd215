"""The f32 variant (dtype="f32"): state, model arithmetic and stored rows in single precision.

BASELINE.json north_star: "paths driven by identical normal draws must match within 1e-12 relative in f64 (f32
variants within a stated tolerance)".  Stated tolerance of this variant against the f64 oracle, on the tested
models: |x_f32 - x_f64| <= 2e-4 * |x_f64| + 2e-4 * scale  (scale = a typical magnitude of the process; the absolute
part covers values that pass near zero, e.g. the Heston variance).  Single-precision rounding of a 250-1000 step
recursion: measured 1e-6 .. 3e-5.
"""
import numpy as np
import pytest
import torch

from conftest import GBM_EQ, HESTON_EQ, grid
from test_gpu_paths import JUMP_EQ, JUMP_INIT, _inject

import sde_sim_rs as S

pytestmark = pytest.mark.gpu


def close(got, ref, scale, rtol=2e-4):
    assert got.dtype == np.float32 and got.shape == ref.shape
    err = np.abs(got.astype(np.float64) - ref)
    bound = rtol * np.abs(ref) + rtol * scale
    worst = float(np.max(err / bound))
    assert worst <= 1.0, worst
    return worst


@pytest.mark.parametrize("name,eqs,times,init,N,wiener,scale", [
    ("gbm", GBM_EQ, grid(252), {"X1": 1.0}, 1024, [True], 1.0),
    ("heston", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 256, [True, True], np.array([100.0, 0.04])),
    ("jump-alg", JUMP_EQ, grid(50, 40), JUMP_INIT, 300, [True, True, False], np.array([0.5, 100.0, 5.0])),
])
@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_f32_identical_draws_within_stated_tolerance(oracle, name, eqs, times, init, N, wiener, scale, scheme):
    U = oracle.Universe(eqs, times)
    inj = _inject(oracle, U, N, "pseudo", 42, wiener)
    ref = oracle.simulate(U, init, N, scheme, inject=inj)
    plan = S.Plan(S.Universe(eqs, times), scheme, "pseudo", inject=torch.from_numpy(inj).cuda(), arithmetic="fast", dtype="f32")
    got = plan.run(init, N).cpu().numpy()
    if name == "jump-alg":
        # a Poisson count is an integer inversion of (u, lambda(t, X)): a last-bit change of X can move a jump; compare the
        # paths whose jump counts agree (all but a handful) — the stated tolerance is about rounding, not about jump times
        same = np.all(np.abs(got.astype(np.float64) - ref)[:, :, 1] < 0.25, axis=1)
        assert same.mean() > 0.97
        got, ref = got[same], ref[same]
    print(name, scheme, "worst err / bound", close(got, ref, scale))


@pytest.mark.parametrize("rng_method,scramble,icdf", [("sobol", "xor", "single"), ("sobol", "xor", "fast"), ("sobol", "none", "reference"),
                                                      ("pseudo", "cp_shift_per_path", "single"), ("sobol", "cp_shift_per_path", "fast")])
def test_f32_gbm_end_to_end(oracle, rng_method, scramble, icdf):
    times, init, N = grid(252), {"X1": 1.0}, 3001
    ref = oracle.simulate(oracle.Universe(GBM_EQ, times), init, N, "euler", rng_method, seed=9, scramble=scramble)
    f = S.simulate(GBM_EQ, times, N, init, rng_method, "euler", seed=9, scramble=scramble, icdf=icdf, arithmetic="fast", dtype="f32")
    got = f.to_numpy()
    ok = np.isfinite(ref).all(axis=(1, 2))                  # unscrambled Sobol: the first point has u = 0 -> NaN path
    assert np.array_equal(ok, np.isfinite(got).all(axis=(1, 2)))
    close(got[ok], ref[ok], 1.0)


def test_f32_layouts_outputs_and_host_run(oracle):
    times, init, N = grid(252, 61), {"S": 100.0, "v": 0.04}, 2500
    U = S.Universe(HESTON_EQ, times)
    kw = dict(scramble="xor", icdf="fast", arithmetic="fast", dtype="f32")
    ref = oracle.simulate(oracle.Universe(HESTON_EQ, times), init, N, "runge-kutta", "sobol", seed=4, scramble="xor")
    scale = np.array([100.0, 0.04])
    ntp = S.Plan(U, "runge-kutta", "sobol", **kw).run(init, N, seed=4).cpu().numpy()
    close(ntp, ref, scale)
    tpn = S.Plan(U, "runge-kutta", "sobol", layout="TPN", **kw).run(init, N, seed=4).cpu().numpy()
    assert tpn.shape == (62, 2, N) and np.array_equal(np.transpose(tpn, (2, 0, 1)), ntp)
    term = S.Plan(U, "runge-kutta", "sobol", output="terminal", **kw).run(init, N, seed=4).cpu().numpy()
    assert term.dtype == np.float32 and np.array_equal(term, ntp[:, -1, :])
    mom = S.Plan(U, "runge-kutta", "sobol", output="moments", **kw).run(init, N, seed=4).cpu().numpy()
    assert mom.dtype == np.float64 and mom.shape == (2, 3)
    t64 = term.astype(np.float64)
    assert np.allclose(mom[:, 0], N) and np.allclose(mom[:, 1], t64.mean(axis=0), rtol=1e-12)
    assert np.allclose(mom[:, 2], ((t64 - t64.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-9)
    host = S.Plan(U, "runge-kutta", "sobol", **kw).run_host(init, N, seed=4)
    assert host.dtype == np.float32 and np.array_equal(host, ntp)


def test_f32_needs_fast_arithmetic():
    with pytest.raises(ValueError, match="f32 needs arithmetic"):
        S.Plan(S.Universe(GBM_EQ, grid(252, 8)), "euler", "pseudo", dtype="f32")
    with pytest.raises(ValueError, match="dtype"):
        S.Plan(S.Universe(GBM_EQ, grid(252, 8)), "euler", "pseudo", dtype="f16", arithmetic="fast")
