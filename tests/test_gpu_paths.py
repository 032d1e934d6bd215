"""GPU parity of the fused simulation kernel against the CPU oracle, through the reference-shaped API.

Bars (BASELINE.json north_star): identical normal draws -> paths within 1e-12 relative (f64);
pseudo-random paths: the ChaCha8 stream is reproduced bit-exactly, so they are held to the same 1e-12;
statistical agreement is checked on top.  FAST icdf/arithmetic: tolerance stated per test.
"""
import numpy as np
import pytest
import torch

from conftest import GBM_EQ, HESTON_EQ, grid

import sde_sim_rs as S

pytestmark = pytest.mark.gpu

JUMP_EQ = ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
           "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
           "C = max(X1 - 100.0, 0.0) + X0"]
JUMP_INIT = {"X0": 0.3, "X1": 100.0, "C": 5.0, "ignored": 1.0}


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _inject(oracle, U, N, rng_method, seed, wiener_mask):
    u = oracle.uniforms(U, N, rng_method, seed=seed)
    S_, K = u.shape[1], u.shape[2]
    buf = np.zeros((N, S_, K + 1))
    for k in range(K):
        buf[:, :, k] = oracle.icdf_normal(u[:, :, k].ravel()).reshape(N, S_) if wiener_mask[k] else u[:, :, k]
    buf[:, :, K] = u[:, :, 0]
    return buf


@pytest.mark.parametrize("name,eqs,times,init,N,wiener", [
    ("C1-gbm", GBM_EQ, grid(252), {"X1": 1.0}, 1024, [True]),
    ("C3-heston", HESTON_EQ, grid(1000), {"S": 100.0, "v": 0.04}, 256, [True, True]),
    ("jump-alg", JUMP_EQ, grid(50, 40), JUMP_INIT, 300, [True, True, False]),
])
@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_identical_draws_paths_within_1e12(oracle, name, eqs, times, init, N, wiener, scheme):
    U = oracle.Universe(eqs, times)
    inj = _inject(oracle, U, N, "pseudo", 42, wiener)
    ref = oracle.simulate(U, init, N, scheme, inject=inj)
    plan = S.Plan(S.Universe(eqs, times), scheme, "pseudo", inject=torch.from_numpy(inj).cuda())
    got = plan.run(init, N).cpu().numpy()
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
@pytest.mark.parametrize("rng_method,scramble", [("pseudo", "cp_shift_per_path"), ("sobol", "cp_shift_per_path"),
                                                 ("sobol", "xor"), ("sobol", "none")])
def test_gbm_end_to_end_matches_oracle(oracle, scheme, rng_method, scramble):
    times, init, N = grid(252), {"X1": 1.0}, 3000          # not a multiple of the CTA size; first CTA has 5 pad lanes
    ref = oracle.simulate(oracle.Universe(GBM_EQ, times), init, N, scheme, rng_method, seed=42, scramble=scramble)
    got = S.simulate(GBM_EQ, times, N, init, rng_method, scheme, seed=42, scramble=scramble).to_numpy()
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


def test_rk_stale_cache_is_reproduced_not_textbook(oracle):
    times, init = grid(252, 3), {"X1": 1.0}
    got = S.simulate(GBM_EQ, times, 1, init, "pseudo", "runge-kutta", seed=0).to_numpy()[0, :, 0]
    assert np.allclose(got, [1.0, 1.0036832451791444, 1.0033430396186047, 1.0068443835552638], rtol=1e-13)   # SURVEY §A.4 trace
    tb = S.simulate(GBM_EQ, times, 1, init, "pseudo", "runge-kutta", seed=0, rk_variant="textbook").to_numpy()[0, :, 0]
    assert np.allclose(tb, [1.0, 1.0036832451791444, 1.0033440972948497, 1.006856639483425], rtol=1e-13)


@pytest.mark.parametrize("rng_method,scramble", [("pseudo", "cp_shift_per_path"), ("sobol", "xor"), ("sobol", "cp_shift_per_path")])
def test_heston_rk_end_to_end_matches_oracle(oracle, rng_method, scramble):
    times, init, N = grid(1000, 200), {"S": 100.0, "v": 0.04}, 500
    ref = oracle.simulate(oracle.Universe(HESTON_EQ, times), init, N, "runge-kutta", rng_method, seed=7, scramble=scramble)
    got = S.simulate(HESTON_EQ, times, N, init, rng_method, "runge-kutta", seed=7, scramble=scramble).to_numpy()
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_jump_model_with_algebraic_end_to_end(oracle, scheme):
    times, N = grid(50, 40), 700
    ref = oracle.simulate(oracle.Universe(JUMP_EQ, times), JUMP_INIT, N, scheme, "pseudo", seed=3)
    got = S.simulate(JUMP_EQ, times, N, JUMP_INIT, "pseudo", scheme, seed=3).to_numpy()
    # a Poisson count can flip where u sits within an ulp of a CDF step (CUDA exp vs glibc exp): compare path-wise
    per_path = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300), axis=(1, 2))
    assert np.mean(per_path <= 1e-12) >= 0.995, np.sort(per_path)[-5:]
    assert got[0, 0, 2] == 5.0                              # algebraic process keeps its initial value at t0


def test_algebraic_stale_cache_quirk_on_device():
    out = S.simulate(["dX = ( A ) * dt", "A = 2.0 + 0.0 * X"], [0.0, 1.0, 2.0, 3.0], 2, {"X": 0.0, "A": 10.0}, "pseudo", "euler",
                     seed=0).to_numpy()[0]
    assert out[:, 0].tolist() == [0.0, 10.0, 10.0, 10.0] and out[:, 1].tolist() == [10.0, 2.0, 2.0, 2.0]


def test_fast_icdf_and_fma_arithmetic_tolerance(oracle):
    # Stated tolerance for icdf="fast" + arithmetic="fast" end to end on C2's model: 1e-11 relative
    times, init, N = grid(252), {"X1": 1.0}, 4096
    ref = oracle.simulate(oracle.Universe(GBM_EQ, times), init, N, "euler", "sobol", seed=42, scramble="xor")
    got = S.simulate(GBM_EQ, times, N, init, "sobol", "euler", seed=42, scramble="xor", icdf="fast", arithmetic="fast").to_numpy()
    e = rel_err(got, ref)
    print("fast end-to-end max rel err", e)
    assert e <= 1e-11


def test_single_precision_icdf_tier_end_to_end(oracle):
    # f64 paths driven by FP32 normals: the path error is the icdf error propagated (sigma sqrt(dt) |dz| per step),
    # orders above 1e-12 by construction and far below the sampling error; all three Sobol kernels must agree on it
    times, N = grid(252), 20_000
    ref = oracle.simulate(oracle.Universe(GBM_EQ, times), {"X1": 1.0}, N, "euler", "sobol", seed=3, scramble="xor")
    outs = []
    for direct in (1, 2, 3):
        got = S.Plan(S.Universe(GBM_EQ, times), "euler", "sobol", scramble="xor", icdf="single", arithmetic="fast",
                     ntp_direct=direct).run({"X1": 1.0}, N, seed=3).cpu().numpy()
        outs.append(got)
        e = rel_err(got, ref)
        assert 1e-12 < e <= 2e-6, e
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])
    for rng_method in ("pseudo",):
        ref = oracle.simulate(oracle.Universe(GBM_EQ, times), {"X1": 1.0}, 2000, "euler", rng_method, seed=3)
        got = S.simulate(GBM_EQ, times, 2000, {"X1": 1.0}, rng_method, "euler", seed=3, icdf="single").to_numpy()
        assert rel_err(got, ref) <= 2e-6


def test_shard_union_is_bit_identical():
    times, init, N = grid(252, 64), {"X1": 1.0}, 5000
    for rng_method, scramble in [("sobol", "xor"), ("sobol", "cp_shift_per_path"), ("pseudo", "none")]:
        whole = S.simulate(GBM_EQ, times, N, init, rng_method, "euler", seed=9, scramble=scramble).to_numpy()
        parts = []
        for r in range(3):
            lo, hi = S.shard_range(N, r, 3)
            parts.append(S.simulate(GBM_EQ, times, hi - lo, init, rng_method, "euler", seed=9, scramble=scramble,
                                    scenario_offset=lo).to_numpy())
        assert np.array_equal(np.concatenate(parts), whole)


def test_layouts_and_reductions_agree():
    times, init, N = grid(1000, 100), {"S": 100.0, "v": 0.04}, 3333
    kw = dict(seed=5, scramble="xor")
    ntp = S.simulate(HESTON_EQ, times, N, init, "sobol", "euler", **kw).to_numpy()
    tpn = S.simulate(HESTON_EQ, times, N, init, "sobol", "euler", layout="TPN", **kw).to_numpy()
    assert np.array_equal(tpn.transpose(2, 0, 1), ntp)
    term = S.simulate(HESTON_EQ, times, N, init, "sobol", "euler", output="terminal", **kw).to_numpy()
    assert np.array_equal(term, ntp[:, -1, :])
    mom = S.simulate(HESTON_EQ, times, N, init, "sobol", "euler", output="moments", **kw).to_numpy()
    assert np.array_equal(mom[:, 0], [N, N])
    assert np.allclose(mom[:, 1], term.mean(axis=0), rtol=1e-13)
    assert np.allclose(mom[:, 2], ((term - term.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-10)


def test_long_format_columns_match_reference_layout():
    times, N = grid(252, 3), 4
    f = S.simulate(HESTON_EQ, times, N, {"S": 100.0, "v": 0.04}, "pseudo", "euler", seed=1)
    c = f.columns()                                         # filtration.rs:108-113
    assert c["scenario"].dtype == np.int32 and c["value"].dtype == np.float64
    assert c["scenario"].tolist()[:8] == [0] * 8 and c["scenario"][-1] == 3
    assert c["process_name"][:4].tolist() == ["S", "v", "S", "v"]
    assert c["time"][:4].tolist() == [times[0], times[0], times[1], times[1]]
    assert c["value"][:2].tolist() == [100.0, 0.04]
    df = f.to_pandas()
    assert list(df.columns) == ["scenario", "time", "process_name", "value"] and len(df) == N * 4 * 2
    tb = f.to_arrow()
    assert [str(t) for t in tb.schema.types] == ["int32", "double", "string", "double"] and tb.num_rows == N * 4 * 2
    assert tb.column("process_name")[1].as_py() == "v" and tb.column("value")[0].as_py() == 100.0


def test_pseudo_mc_statistics_match_closed_form():
    # Euler-GBM: E[X_S] = (1 + mu dt)^S; Var from E[X^2] = ((1+mu dt)^2 + sigma^2 dt)^S (unit-variance normals;
    # the A&S map perturbs Var(z) at ~1e-4, SURVEY §B.4 — CIs below are wider than that)
    D, N = 252, 1 << 20
    m = S.simulate(GBM_EQ, grid(D), N, {"X1": 1.0}, "pseudo", "euler", seed=123, output="moments").moments()["X1"]
    mu, sig, dt = 0.05, 0.1, 1.0 / D
    mean = (1 + mu * dt) ** D
    var = ((1 + mu * dt) ** 2 + sig * sig * dt) ** D - mean**2
    se = np.sqrt(var / N)
    assert abs(m["mean"] - mean) < 5 * se
    assert abs(m["variance"] / var - 1) < 5 * np.sqrt(2.0 / N) + 1e-3


def test_rqmc_beats_mc_on_the_mean():
    # one XOR mask per run keeps the net structure: integration error of E[X_T] far below the MC standard error
    D, N = 64, 1 << 16
    mean = (1 + 0.05 / D) ** D
    errs = []
    for seed in range(4):
        m = S.simulate(GBM_EQ, grid(D), N, {"X1": 1.0}, "sobol", "euler", seed=seed, scramble="xor", output="moments").moments()["X1"]
        errs.append(abs(m["mean"] - mean))
    mc_se = 0.1 / np.sqrt(N)
    assert np.mean(errs) < 0.2 * mc_se


def test_run_host_matches_device_run():
    times, init, N = grid(252, 50), {"X1": 1.0}, 10_000
    plan = S.Plan(S.Universe(GBM_EQ, times), "euler", "sobol", scramble="xor")
    dev = plan.run(init, N, seed=4).cpu().numpy()
    host = plan.run_host(init, N, seed=4)
    assert np.array_equal(host, dev)
    pinned = torch.empty((N, 51, 1), dtype=torch.float64).pin_memory()
    plan.run_host(init, N, seed=4, out=pinned)
    assert np.array_equal(pinned.numpy(), dev)


# ---------------------------------------------------------------- remaining BASELINE configs + store paths
def test_direct_and_transposed_store_paths_are_bit_identical():
    # full-path [N][T][P] output has two implementations: 256-bit sector stores straight from
    # registers vs the shared-memory transpose; same arithmetic, so the bytes must agree
    for eqs, times, init, N in [(GBM_EQ, grid(252), {"X1": 1.0}, 5003), (HESTON_EQ, grid(1000, 123), {"S": 100.0, "v": 0.04}, 2050)]:
        outs = []
        for direct in (1, 2, 3):
            if direct == 3 and len(times) > 600:
                continue                                    # (T-1) K 128 B of tables must fit in shared memory
            plan = S.Plan(S.Universe(eqs, times), "euler", "sobol", scramble="xor", ntp_direct=direct)
            outs.append(plan.run(init, N, seed=11, scenario_offset=77).cpu().numpy())
        assert all(np.array_equal(outs[0], o) for o in outs[1:])      # reference icdf, strict arithmetic: same operations


@pytest.mark.parametrize("scramble", ["xor", "none"])
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_persistent_kernel_many_items_per_warp_matches_tiled(scramble, mode):
    # sde_sim_resident.cuh: every warp of the resident CTAs walks several 32-path items — including pad lanes at both
    # ends and a scenario offset.  With the reference inverse normal and strict arithmetic the two kernels execute
    # the same operations, so the bytes must agree; the fast inverse normal of the persistent kernel forms the
    # exponent term arithmetically (one rounding less), so there the bound is the stated icdf tolerance
    kw = dict(icdf="reference", arithmetic="strict") if mode == "strict" else dict(icdf="fast", arithmetic="fast")
    for eqs, init, steps, N, off in ((GBM_EQ, {"X1": 1.0}, 13, 700_001, 123_457), (HESTON_EQ, {"S": 100.0, "v": 0.04}, 9, 333_333, 2)):
        times = grid(252, steps)
        outs = []
        for direct in (2, 3):
            plan = S.Plan(S.Universe(eqs, times), "euler", "sobol", scramble=scramble, ntp_direct=direct, **kw)
            assert ("sde_sim_resident.cuh" in plan.source) == (direct == 3)
            outs.append(plan.run(init, N, seed=5, scenario_offset=off).cpu().numpy())
        if mode == "strict":
            assert np.array_equal(outs[0], outs[1], equal_nan=True)
        else:
            ok = np.isfinite(outs[0])
            assert np.array_equal(ok, np.isfinite(outs[1]))
            assert rel_err(outs[1][ok], outs[0][ok]) <= 1e-12


@pytest.mark.parametrize("rng_method,scramble,scheme,steps,N", [("sobol", "xor", "runge-kutta", 1000, 600), ("pseudo", "cp_shift_per_path", "euler", 37, 300),
                                                            ("sobol", "cp_shift_per_path", "euler", 50, 257)])
@pytest.mark.parametrize("mode", [4, 5])
def test_bulk_copy_store_path_is_bit_identical(rng_method, scramble, scheme, steps, N, mode):
    # ntp_direct=4: rows staged per lane in shared memory and written by cp.async.bulk (P even); ntp_direct=5: one 2-D
    # tensor-map store per warp and 128-byte box row (P = 2 or 4); same values as the default path
    if mode == 5 and rng_method != "sobol":
        pytest.skip("tensor-map stores ride on 4-step groups (K <= 2, no ChaCha block alignment)")
    times, init = grid(1000, steps), {"S": 100.0, "v": 0.04}
    kw = dict(scramble=scramble, seed=9, scenario_offset=3)
    a = S.Plan(S.Universe(HESTON_EQ, times), scheme, rng_method, scramble=scramble, ntp_direct=mode)
    assert f"#define SDE_TMA {mode - 3}" in a.source
    got = a.run(init, N, seed=9, scenario_offset=3).cpu().numpy()
    ref = S.simulate(HESTON_EQ, times, N, init, rng_method, scheme, **kw).to_numpy()
    assert np.array_equal(got, ref)
    with pytest.raises(ValueError, match="even number of processes"):
        S.Plan(S.Universe(GBM_EQ, times), "euler", "pseudo", ntp_direct=4)


@pytest.mark.parametrize("scheme,scramble,kw", [("runge-kutta", "xor", dict(icdf="fast", arithmetic="fast")), ("euler", "xor", dict(icdf="fast", arithmetic="fast")),
                                                 ("runge-kutta", "none", dict())])
def test_persistent_kernel_with_global_lane_table_is_bit_identical(scheme, scramble, kw):
    # ntp_direct=3 on a time grid whose lane table does not fit in shared memory (2 x 1000 dimensions): the table is prepared
    # by the host in global memory (per seed); same values as the time-tiled kernel the lowering picks on its own
    times, init, N = grid(1000), {"S": 100.0, "v": 0.04}, 700
    forced = S.Plan(S.Universe(HESTON_EQ, times), scheme, "sobol", scramble=scramble, ntp_direct=3, **kw)
    assert "#define SDE_RES_LANE_GLOBAL 1" in forced.source
    auto = S.Plan(S.Universe(HESTON_EQ, times), scheme, "sobol", scramble=scramble, **kw)
    assert "sde_sim_kernel.cuh" in auto.source
    for seed, off in ((9, 3), (10, 251)):                  # a second seed: the prepared table follows the masks
        a = forced.run(init, N, seed=seed, scenario_offset=off)
        b = auto.run(init, N, seed=seed, scenario_offset=off)
        assert torch.equal(a, b)


def test_direct_store_needs_aligned_output():
    plan = S.Plan(S.Universe(GBM_EQ, grid(252, 8)), "euler", "sobol", scramble="xor")
    buf = torch.empty(100 * 9 + 1, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError, match="32-byte aligned"):
        plan.run({"X1": 1.0}, 100, out=buf[1:].view(100, 9, 1))


@pytest.mark.parametrize("eqs,times,init,N,wiener,scheme", [
    (GBM_EQ, grid(252), {"X1": 1.0}, 1024, [True], "euler"),
    (HESTON_EQ, grid(1000, 300), {"S": 100.0, "v": 0.04}, 256, [True, True], "runge-kutta"),
])
def test_fast_arithmetic_identical_draws_within_1e12(oracle, eqs, times, init, N, wiener, scheme):
    # arithmetic="fast" (FMA contraction + multiplicative rewrite of a_j*X coefficients): identical normal draws
    # must still reproduce the oracle's strictly ordered f64 arithmetic to 1e-12 relative
    U = oracle.Universe(eqs, times)
    inj = _inject(oracle, U, N, "pseudo", 5, wiener)
    ref = oracle.simulate(U, init, N, scheme, inject=inj)
    plan = S.Plan(S.Universe(eqs, times), scheme, "pseudo", inject=torch.from_numpy(inj).cuda(), arithmetic="fast")
    got = plan.run(init, N).cpu().numpy()
    assert rel_err(got, ref) <= 1e-12, rel_err(got, ref)


@pytest.mark.parametrize("arithmetic,icdf,tol", [("strict", "reference", 1e-12), ("fast", "fast", 1e-12)])
def test_c4_basket_64_factors_matches_oracle(oracle, arithmetic, icdf, tol):
    # C4: 64-dim correlated GBM basket, Sobol dims = 64 x 252 = 16128, correlation through shared dW names
    from conftest import basket_equations

    eqs, init = basket_equations(64)
    times, N = grid(252), 96
    U = oracle.Universe(eqs, times)
    assert (U.P, U.K) == (64, 64) and U.factors[:3] == ["dW1", "dW2", "dW3"]
    ref = oracle.simulate(U, init, N, "euler", "sobol", seed=3, scramble="xor")
    got = S.simulate(eqs, times, N, init, "sobol", "euler", seed=3, scramble="xor", arithmetic=arithmetic, icdf=icdf).to_numpy()
    assert rel_err(got, ref) <= tol, rel_err(got, ref)
    mom = S.simulate(eqs, times, N, init, "sobol", "euler", seed=3, scramble="xor", arithmetic=arithmetic, icdf=icdf,
                     output="moments").to_numpy()
    assert np.allclose(mom[:, 1], got[:, -1, :].mean(axis=0), rtol=1e-13)


def test_c5_terminal_moments_pseudo_match_oracle(oracle):
    # C5 shape: GBM, 365 steps, pseudo-random, terminal moments only (reduced N for the CPU oracle)
    times, N = grid(365), 1 << 13
    ref = oracle.simulate(oracle.Universe(GBM_EQ, times), {"X1": 1.0}, N, "euler", "pseudo", seed=77)[:, -1, 0]
    for icdf, arithmetic, tol in (("reference", "strict", 1e-12), ("fast", "fast", 1e-12)):
        m = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "pseudo", "euler", seed=77, output="moments", icdf=icdf,
                       arithmetic=arithmetic).to_numpy()[0]
        assert m[0] == N
        assert abs(m[1] / ref.mean() - 1) <= tol
        assert abs(m[2] / ((ref - ref.mean()) ** 2).sum() - 1) <= 1e-9
        term = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "pseudo", "euler", seed=77, output="terminal", icdf=icdf,
                          arithmetic=arithmetic).to_numpy()[:, 0]
        assert rel_err(term, ref) <= tol


def test_full_size_c2_properties():
    # C2 at BASELINE size (2^24 paths x 252 steps, 34 GB): size-independent properties instead of an oracle run
    N, D = 1 << 24, 252
    plan = S.Plan(S.Universe(GBM_EQ, grid(D)), "euler", "sobol", scramble="xor", icdf="fast", arithmetic="fast")
    out = plan.run({"X1": 1.0}, N, seed=42)
    assert bool((out[:, 0, 0] == 1.0).all())                                 # t0 row = initial value for every path
    assert bool(torch.isfinite(out).all()) and bool((out > 0).all())
    term = out[:, -1, 0]
    mean = (1 + 0.05 / D) ** D
    # a 2^24-point scrambled net integrates E[X_T] far below the MC standard error 0.1/sqrt(N) = 2.4e-5
    assert abs(float(term.mean()) - mean) < 3e-6
    # shard check at full size: paths [2^23, 2^23 + 4096) recomputed with an offset are bit-identical
    lo = 1 << 23
    part = plan.run({"X1": 1.0}, 4096, seed=42, scenario_offset=lo)
    assert torch.equal(part, out[lo:lo + 4096])
    # dimension-wise net property: in every time step the 2^24 uniforms hit each of 2^12 equal cells 2^12 times,
    # hence each step's log-increment sample mean is tiny
    inc = torch.log(out[:, 1:9, 0] / out[:, 0:8, 0]).mean(dim=0)
    assert bool((inc.abs() < 2e-4).all())
    del out
    torch.cuda.empty_cache()


def test_full_size_c3_properties(oracle):
    # C3 at BASELINE size (Heston, Runge-Kutta, 2^22 paths x 1000 steps, P = 2: 67 GB): the oracle on the first and last paths
    # of the buffer, and size-independent properties for the rest
    N, D = 1 << 22, 1000
    init = {"S": 100.0, "v": 0.04}
    times = grid(D)
    plan = S.Plan(S.Universe(HESTON_EQ, times), "runge-kutta", "sobol", scramble="xor", icdf="fast", arithmetic="fast")
    out = plan.run(init, N, seed=42)
    assert out.shape == (N, D + 1, 2)
    assert bool((out[:, 0, 0] == 100.0).all()) and bool((out[:, 0, 1] == 0.04).all())      # t0 row = initial values
    assert bool(torch.isfinite(out).all()) and bool((out[:, :, 0] > 0).all())
    U = oracle.Universe(HESTON_EQ, times)
    head = oracle.simulate(U, init, 48, "runge-kutta", "sobol", seed=42, scramble="xor")
    tail = oracle.simulate(U, init, 48, "runge-kutta", "sobol", seed=42, scramble="xor", scenario_offset=N - 48)
    for got, ref in ((out[:48].cpu().numpy(), head), (out[N - 48:].cpu().numpy(), tail)):
        scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=(0, 1), keepdims=True) * 1e-3)   # v can sit near zero
        assert np.max(np.abs(got - ref) / scale) <= 1e-10                      # fast tier: 5e-13 per draw, 1000 steps, sqrt(v) near 0
    # shard check at full size: paths around the middle recomputed with an offset are bit-identical
    lo = (1 << 21) - 300
    part = plan.run(init, 777, seed=42, scenario_offset=lo)
    assert torch.equal(part, out[lo:lo + 777])
    # the moments-only plan (another store path and the warp-shuffle / block reduction) sees the same terminal values
    mom = S.simulate(HESTON_EQ, times, N, init, "sobol", "runge-kutta", seed=42, scramble="xor", icdf="fast", arithmetic="fast",
                     output="moments").to_numpy()
    term = out[:, -1, :].double()
    assert mom[0, 0] == N and mom[1, 0] == N
    for p in range(2):
        mean = float(term[:, p].mean())
        assert abs(mom[p, 1] / mean - 1) < 1e-11
        assert abs(mom[p, 2] / float(((term[:, p] - mean) ** 2).sum()) - 1) < 1e-8
    # sanity of the model itself: S drifts at 5 % a year (the reference's stale-cache Runge-Kutta variant is not the
    # textbook scheme: a few 1e-2 relative is all that is claimed), v stays around its mean-reversion level
    assert abs(mom[0, 1] / (100.0 * np.exp(0.05)) - 1) < 5e-2 and abs(mom[1, 1] - 0.04) < 1e-2
    del out
    torch.cuda.empty_cache()


def test_full_size_c5_per_gpu_moments_against_closed_form():
    # C5's per-GPU share (2^30 paths x 365 steps, pseudo-random ChaCha8 streams, moments only): closed-form Euler-GBM moments
    N, D = 1 << 30, 365
    m = S.simulate(GBM_EQ, grid(D), N, {"X1": 1.0}, "pseudo", "euler", seed=2024, output="moments", icdf="fast",
                   arithmetic="fast").to_numpy()[0]
    assert m[0] == N
    mu, sig, dt = 0.05, 0.1, 1.0 / D
    mean = (1 + mu * dt) ** D
    var = ((1 + mu * dt) ** 2 + sig * sig * dt) ** D - mean**2
    assert abs(m[1] - mean) < 5 * np.sqrt(var / N)                            # 5 standard errors = 1.6e-5
    assert abs(m[2] / (N - 1) / var - 1) < 1e-3                               # A&S perturbs Var(z) at ~1e-4


# ---------------------------------------------------------------- ragged shapes (tile / group / sector edges)
THREE_EQ = ["dA = ( 0.3 * (1.0 - A) ) * dt + ( 0.2 ) * dW1",
            "dB = ( 0.1 * B ) * dt + ( 0.3 * B ) * dW2",
            "dC = ( A - C ) * dt + ( 0.1 * B ) * dW1"]
THREE_INIT = {"A": 0.5, "B": 2.0, "C": -1.0}


@pytest.mark.parametrize("steps", [1, 2, 3, 4, 5, 7, 8, 35, 36, 37, 40, 73, 145])
def test_ragged_step_counts_all_store_paths(oracle, steps):
    # the direct sector-store path shifts every warp's step groups by gamma in 0..3 and handles the first gamma and
    # the last <= 3 steps on their own; tiles are 32-36 steps: cover every residue, both store paths, odd offsets
    for eqs, init, D in ((GBM_EQ, {"X1": 1.0}, 252), (HESTON_EQ, {"S": 100.0, "v": 0.04}, 1000), (THREE_EQ, THREE_INIT, 50)):
        times, N, off = grid(D, steps), 777, 1021
        ref = oracle.simulate(oracle.Universe(eqs, times), init, N, "euler", "sobol", seed=21, scramble="xor", scenario_offset=off)
        for direct in (1, 2, 3):
            plan = S.Plan(S.Universe(eqs, times), "euler", "sobol", scramble="xor", ntp_direct=direct)
            got = plan.run(init, N, seed=21, scenario_offset=off).cpu().numpy()
            assert rel_err(got, ref) <= 1e-12, (steps, direct, len(eqs), rel_err(got, ref))


@pytest.mark.parametrize("steps", [33, 34, 65, 66, 97, 98, 130])
def test_last_tile_of_one_or_two_steps(oracle, steps):
    # time-tiled kernel with direct sector stores: a last tile of 1-2 steps while a warp's last full group ended before the
    # tile boundary (S mod TT in {1, 2}; tile_steps=32 pins TT): the <= 3 trailing steps must read the tile that stages them.
    # Both users of that kernel: Sobol plans (ntp_direct=2) and injected draws; scenario offsets cover all four gammas.
    times = grid(252, steps)
    U = oracle.Universe(GBM_EQ, times)
    for off in (0, 1, 2, 3):
        ref = oracle.simulate(U, {"X1": 1.0}, 700, "euler", "sobol", seed=3, scramble="xor", scenario_offset=off)
        for tt in (0, 32):
            plan = S.Plan(S.Universe(GBM_EQ, times), "euler", "sobol", scramble="xor", ntp_direct=2, tile_steps=tt)
            got = plan.run({"X1": 1.0}, 700, seed=3, scenario_offset=off).cpu().numpy()
            assert rel_err(got, ref) <= 1e-12, (steps, off, tt, rel_err(got, ref))
    inj = _inject(oracle, U, 700, "pseudo", 11, [True])
    ref = oracle.simulate(U, {"X1": 1.0}, 700, "euler", inject=inj)
    for tt in (0, 32):
        plan = S.Plan(S.Universe(GBM_EQ, times), "euler", "pseudo", inject=torch.from_numpy(inj).cuda(), tile_steps=tt)
        assert rel_err(plan.run({"X1": 1.0}, 700).cpu().numpy(), ref) <= 1e-12, (steps, tt)


@pytest.mark.parametrize("N,off", [(1, 0), (2, 3), (5, 250), (251, 5), (256, 251), (257, 0), (1000, 4091)])
def test_ragged_scenario_counts_and_offsets(oracle, N, off):
    times = grid(252, 41)
    for rng_method, scramble in (("sobol", "xor"), ("pseudo", "none")):
        ref = oracle.simulate(oracle.Universe(GBM_EQ, times), {"X1": 1.0}, N, "runge-kutta", rng_method, seed=8,
                              scramble=scramble if rng_method == "sobol" else "cp_shift_per_path", scenario_offset=off)
        got = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, rng_method, "runge-kutta", seed=8,
                         scramble=scramble if rng_method == "sobol" else "cp_shift_per_path", scenario_offset=off).to_numpy()
        assert rel_err(got, ref) <= 1e-12


def test_pure_drift_and_zero_term_processes(oracle):
    # K = 0 (no stochastic factor), a zero-term SDE (`delta = ...` quirk, util.rs:80-82) and an algebraic process
    eqs = ["dX = ( 0.5 * X ) * dt", "delta = 1.0", "Y = X * 2.0 + t"]
    times, init = grid(10, 12), {"X": 1.0, "elta": 3.0, "Y": 7.0}
    ref = oracle.simulate(oracle.Universe(eqs, times), init, 9, "euler", "pseudo", seed=1)
    got = S.simulate(eqs, times, 9, init, "pseudo", "euler", seed=1).to_numpy()
    assert rel_err(got[:, :, [0, 2]], ref[:, :, [0, 2]]) <= 1e-12 and np.array_equal(got[:, :, 1], ref[:, :, 1])
    with pytest.raises(ValueError, match="runge-kutta needs"):
        S.simulate(eqs, times, 9, init, "pseudo", "runge-kutta", seed=1)


def test_nonuniform_time_grid(oracle):
    times = [0.0, 0.01, 0.015, 0.1, 0.1000001, 0.5, 2.0, 2.5]
    for scheme in ("euler", "runge-kutta"):
        ref = oracle.simulate(oracle.Universe(HESTON_EQ, times), {"S": 100.0, "v": 0.04}, 300, scheme, "sobol", seed=2, scramble="xor")
        got = S.simulate(HESTON_EQ, times, 300, {"S": 100.0, "v": 0.04}, "sobol", scheme, seed=2, scramble="xor").to_numpy()
        assert rel_err(got, ref) <= 1e-12


def test_run_host_chunk_boundaries():
    # sde_plan_run_host cuts the scenarios into ~512 MiB chunks (double-buffered D2H); chunk seams must not show
    times, init = grid(252), {"X1": 1.0}
    N = 600_011                                             # > 2 chunks of 265 216 paths for [N, 253, 1]
    plan = S.Plan(S.Universe(GBM_EQ, times), "euler", "sobol", scramble="xor", icdf="fast", arithmetic="fast")
    dev = plan.run(init, N, seed=4, scenario_offset=123)
    host = torch.empty((N, 253, 1), dtype=torch.float64).pin_memory()
    plan.run_host(init, N, seed=4, scenario_offset=123, out=host)
    assert torch.equal(host, dev.cpu())
    term_plan = S.Plan(S.Universe(GBM_EQ, times), "euler", "pseudo", output="terminal")
    assert np.array_equal(term_plan.run_host(init, 70_001, seed=4), term_plan.run(init, 70_001, seed=4).cpu().numpy())


def test_simulate_frame_returns_the_reference_long_frame():
    # the reference's call and return value (src/py_binding.rs:10-55): long frame, (scenario, time, process) row order
    times, N = grid(252, 5), 7
    df = S.simulate_frame(HESTON_EQ, times, N, {"S": 100.0, "v": 0.04}, "sobol", "euler", seed=3)
    assert list(df.columns) == ["scenario", "time", "process_name", "value"] and len(df) == N * 6 * 2
    dense = S.simulate(HESTON_EQ, times, N, {"S": 100.0, "v": 0.04}, "sobol", "euler", seed=3).to_numpy()
    assert np.array_equal(np.asarray(df["value"], dtype=np.float64), dense.reshape(-1))
    assert [str(x) for x in list(df["process_name"][:3])] == ["S", "v", "S"] and int(np.asarray(df["scenario"])[-1]) == N - 1


@pytest.mark.parametrize("script", ["example_gbm.py", "example_jumps.py"])
def test_examples_run(script):
    import os
    import subprocess
    import sys

    from conftest import PKG, ROOT

    env = dict(os.environ, PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "scenario" in r.stdout and "process_name" in r.stdout


def test_simulate_devices_union_equals_single_run():
    # one process driving several GPUs: shards by scenario offset, union bit-identical to one run; moments Chan-merged.
    # On a one-GPU box the same device is listed three times (three shards, same stream) — the sharding logic is what is tested.
    devs = list(range(torch.cuda.device_count()))
    devs = devs if len(devs) > 1 else [0, 0, 0]
    times, N = grid(252, 30), 1001
    kw = dict(seed=17, scramble="xor", icdf="fast", arithmetic="fast")
    whole = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "sobol", "euler", **kw).to_numpy()
    shards = S.simulate_devices(GBM_EQ, times, N, {"X1": 1.0}, "sobol", "euler", devices=devs, **kw)
    assert [s.scenario_offset for s in shards] == [S.shard_range(N, i, len(devs))[0] for i in range(len(devs))]
    assert np.array_equal(np.concatenate([s.to_numpy() for s in shards]), whole)
    m = S.simulate_devices(GBM_EQ, times, N, {"X1": 1.0}, "sobol", "euler", devices=devs, output="moments", **kw).to_numpy()[0]
    assert m[0] == N and abs(m[1] / whole[:, -1, 0].mean() - 1) <= 1e-13
    tiny = S.simulate_devices(GBM_EQ, times, 2, {"X1": 1.0}, "pseudo", "euler", devices=devs, seed=1)     # fewer scenarios than devices
    assert sum(s.shape[0] for s in tiny) == 2


def test_ahead_of_time_cache_is_hit_by_a_second_process():
    # plan creation NVRTC-compiles once per (model, options, headers, NVRTC version) and leaves the cubin in the on-disk cache
    # (build() fills it for the BASELINE configs); a later process loads it without compiling
    import os
    import subprocess
    import sys

    from conftest import PKG

    code = ("import sde_sim_rs as S; p = S.Plan(S.Universe(['dX1 = ( 0.0123 * X1 ) * dt + ( 0.1717 * X1) * dW1'], "
            "[k / 7 for k in range(8)]), 'euler', 'pseudo'); print('prelowered', p.prelowered)")
    env = dict(os.environ, PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""))
    first = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    second = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert first.returncode == 0 and second.returncode == 0, first.stderr + second.stderr
    assert "prelowered True" in second.stdout


def test_pseudo_inverse_normal_rare_path_is_bit_identical(monkeypatch):
    # The ChaCha-driven fast tier evaluates the inverse normal from the two 32-bit words of a draw and redoes the draws with
    # min(p, 1-p) < 2^-32 (one in 2^32) through the general 53-bit entry in a per-group fix-up.  The test hook widens "rare" to
    # min(p, 1-p) < 2^-8, so thousands of draws take the fix-up (block regenerated from the counter, slot picked, draw replaced):
    # the results must not change by a single bit, for one factor, two factors (Heston) and a step count that leaves single steps.
    cases = [(GBM_EQ, {"X1": 1.0}, "euler", grid(365, 37)), (HESTON_EQ, {"S": 100.0, "v": 0.04}, "runge-kutta", grid(1000, 23))]
    for eqs, init, scheme, times in cases:
        kw = dict(icdf="fast", arithmetic="fast")
        monkeypatch.delenv("SDE_B200_DEFINES", raising=False)
        base = S.Plan(S.Universe(eqs, times), scheme, "pseudo", **kw).run(init, 3000, seed=11).cpu().numpy()
        monkeypatch.setenv("SDE_B200_DEFINES", "SDE_W64_RARE_SHIFT=24")
        hooked = S.Plan(S.Universe(eqs, times), scheme, "pseudo", **kw)
        assert "#define SDE_W64_RARE_SHIFT 24" in hooked.source
        assert np.array_equal(hooked.run(init, 3000, seed=11).cpu().numpy(), base)
    monkeypatch.delenv("SDE_B200_DEFINES", raising=False)
