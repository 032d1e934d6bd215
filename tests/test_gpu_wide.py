"""GPU parity of the FP64 tensor-core kernel for wide linear models (sde_sim_wide.cuh) against the CPU oracle.

The oracle restates euler_iteration (src/sim/euler.rs:5-37) term by term; the kernel evaluates the same step as
X_i *= 1 + a_i dt + sqrt(dt) sum_k M[i][k] z_k with the sum on DMMA (mma.sync.m8n8k4.f64).  Tolerance: 1e-12 relative
on terminal values (north_star: paths within 1e-12 in f64), stated in each test."""
import numpy as np
import pytest
import torch

from conftest import basket_equations, grid

import sde_sim_rs as S

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _plan(eqs, times, output, wide_mma, icdf="fast"):
    return S.Plan(S.Universe(eqs, times), "euler", "sobol", output=output, scramble="xor", icdf=icdf, arithmetic="fast",
                  wide_mma=wide_mma)


@pytest.mark.parametrize("icdf", ["fast", "reference"])
def test_c4_basket_terminal_on_tensor_path_matches_oracle(oracle, icdf):
    # C4 shape: 64 assets x 64 factors, 252 steps (Sobol dims 16128); terminal values of 100 paths at a scenario offset
    eqs, init = basket_equations(64)
    times, N, off = grid(252), 100, 37
    plan = _plan(eqs, times, "terminal", 2, icdf)
    assert "sde_sim_wide.cuh" in plan.source
    got = plan.run(init, N, seed=3, scenario_offset=off).cpu().numpy()
    ref = oracle.simulate(oracle.Universe(eqs, times), init, N, "euler", "sobol", seed=3, scramble="xor", scenario_offset=off)[:, -1, :]
    assert got.shape == (N, 64)
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


@pytest.mark.parametrize("n_assets,N,off", [(16, 1, 0), (20, 37, 0), (33, 50, 11), (40, 9, 3), (120, 19, 5)])
def test_ragged_widths_and_path_counts(oracle, n_assets, N, off):
    # P = K not a multiple of 8 / 4 (pad tiles), one / two row tiles per warp, path counts that leave pad rows
    eqs, init = basket_equations(n_assets)
    times = grid(252, 24)
    plan = _plan(eqs, times, "terminal", 2)
    got = plan.run(init, N, seed=11, scenario_offset=off).cpu().numpy()
    ref = oracle.simulate(oracle.Universe(eqs, times), init, N, "euler", "sobol", seed=11, scramble="xor", scenario_offset=off)[:, -1, :]
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


def test_fewer_factors_than_assets_and_nonuniform_grid(oracle):
    # 24 assets loaded on 17 factors (K not a multiple of 4, triangular only in the leading block), irregular time grid
    rng = np.random.default_rng(5)
    P, K = 24, 17
    L = np.abs(rng.normal(size=(P, K))) * 0.05      # literal `c * X` coefficients (a leading minus is a unary node)
    eqs = []
    for i in range(P):
        terms = [f"( {0.01 * (i + 1):.17g} * A{i} ) * dt"] + [f"( {L[i, k]:.17g} * A{i} ) * dW{k + 1}" for k in range(K) if (i + k) % 3]
        eqs.append(f"dA{i} = " + " + ".join(terms))
    init = {f"A{i}": 1.0 + i for i in range(P)}
    times = [0.0, 0.01, 0.03, 0.035, 0.1, 0.11, 0.2, 0.5, 0.51, 1.0]
    N = 70
    plan = _plan(eqs, times, "terminal", 2)
    got = plan.run(init, N, seed=9).cpu().numpy()
    ref = oracle.simulate(oracle.Universe(eqs, times), init, N, "euler", "sobol", seed=9, scramble="xor")[:, -1, :]
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


def test_moments_match_terminal_values_and_tiled_kernel():
    # moments = per-warp Chan merges + sde_moments_finalize; against numpy over the terminal values of the same plan shape,
    # and against the time-tiled kernel (wide_mma = 1: off) on the same inputs
    eqs, init = basket_equations(64)
    times, N = grid(252, 32), 5000
    term = _plan(eqs, times, "terminal", 2).run(init, N, seed=21).cpu().numpy()
    mom = _plan(eqs, times, "moments", 2).run(init, N, seed=21).cpu().numpy()
    assert mom.shape == (64, 3) and bool((mom[:, 0] == N).all())
    assert np.allclose(mom[:, 1], term.mean(axis=0), rtol=1e-13)
    assert np.allclose(mom[:, 2], ((term - term.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-9)
    tiled = _plan(eqs, times, "terminal", 1)
    assert "sde_sim_wide.cuh" not in tiled.source
    assert rel_err(term, tiled.run(init, N, seed=21).cpu().numpy()) <= 1e-12


def test_auto_selection_and_refusal():
    eqs, init = basket_equations(64)
    times = grid(252, 8)
    assert "sde_sim_wide.cuh" in _plan(eqs, times, "moments", 0).source          # auto: qualifies
    assert "sde_sim_wide.cuh" in _plan(eqs, times, "paths", 0).source
    tpn = S.Plan(S.Universe(eqs, times), "euler", "sobol", output="paths", layout="TPN", scramble="xor", icdf="fast", arithmetic="fast")
    assert "sde_sim_wide.cuh" not in tpn.source                                    # the transposed layout stays on the time-tiled kernel
    with pytest.raises(ValueError):                                                # required but not a wide linear model
        S.Plan(S.Universe(["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"], times), "euler", "sobol", output="moments",
               scramble="xor", icdf="fast", arithmetic="fast", wide_mma=2)


@pytest.mark.parametrize("n_assets,N,off", [(64, 40, 0), (20, 37, 6), (33, 21, 0)])
def test_full_paths_in_reference_order(oracle, n_assets, N, off):
    # [N][T][P] rows incl. the t0 row (filtration.rs:87-113): 128-bit row stores (P even, whole tiles) and the scalar tail / odd-P path
    eqs, init = basket_equations(n_assets)
    times = grid(252, 40)
    plan = _plan(eqs, times, "paths", 2)
    got = plan.run(init, N, seed=13, scenario_offset=off).cpu().numpy()
    ref = oracle.simulate(oracle.Universe(eqs, times), init, N, "euler", "sobol", seed=13, scramble="xor", scenario_offset=off)
    assert got.shape == ref.shape == (N, 41, n_assets)
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)
    # same rows through the host-buffer call (chunked launches + D2H)
    host = plan.run_host(init, N, seed=13, scenario_offset=off)
    assert np.array_equal(np.asarray(host).reshape(got.shape), got)


def test_shard_union_equals_single_run():
    # disjoint scenario ranges (multi-GPU sharding): the union of two shards is bit-identical to one run
    eqs, init = basket_equations(32)
    times, N = grid(252, 16), 301
    plan = _plan(eqs, times, "terminal", 2)
    whole = plan.run(init, N, seed=4).cpu().numpy()
    a = plan.run(init, 150, seed=4, scenario_offset=0).cpu().numpy()
    b = plan.run(init, 151, seed=4, scenario_offset=150).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b]), whole)


def test_full_size_c4_moments_against_closed_form():
    # C4 at BASELINE size (2^20 paths x 252 steps, 64 assets, Sobol dims 16128, moments): size-independent properties.
    # Euler-GBM per asset: E[S_T] = S0 (1 + mu dt)^S and E[S_T^2] = S0^2 ((1 + mu dt)^2 + sigma_i^2 dt)^S with
    # sigma_i^2 = sum_k M[i][k]^2 (rows of the Cholesky factor have unit norm, so sigma_i is the asset's volatility).
    eqs, init = basket_equations(64)
    N, D = 1 << 20, 252
    plan = _plan(eqs, grid(D), "moments", 0)
    assert "sde_sim_wide.cuh" in plan.source
    mom = plan.run(init, N, seed=42).cpu().numpy()
    assert bool((mom[:, 0] == N).all())
    mu, dt = 0.05, 1.0 / D
    sig = 0.1 + 0.2 * np.arange(64) / 63
    mean = 100.0 * (1 + mu * dt) ** D
    var = 100.0**2 * ((1 + mu * dt) ** 2 + sig**2 * dt) ** D - mean**2
    se = np.sqrt(var / N)
    # RQMC integrates the mean far below the MC standard error; the A&S map perturbs Var(z) at ~1e-4
    assert bool((np.abs(mom[:, 1] - mean) < 1.0 * se).all()), np.max(np.abs(mom[:, 1] - mean) / se)
    # sample variance of a near-lognormal terminal value (kurtosis < 5 at sigma <= 0.3): relative standard error
    # sqrt((kurt - 1) / N) <= 2e-3; 5 of those + the A&S term
    assert bool((np.abs(mom[:, 2] / (N - 1) / var - 1) < 5 * 2e-3 + 1e-3).all()), np.max(np.abs(mom[:, 2] / (N - 1) / var - 1))
    # two half-size shards merged with the library's Chan merge reproduce the single run to rounding
    a = plan.run(init, N // 2, seed=42, scenario_offset=0).cpu().numpy()
    b = plan.run(init, N // 2, seed=42, scenario_offset=N // 2).cpu().numpy()
    merged = S.merge_moments(np.stack([a, b]))
    assert np.allclose(merged[:, 1], mom[:, 1], rtol=1e-13) and np.allclose(merged[:, 2], mom[:, 2], rtol=1e-10)
