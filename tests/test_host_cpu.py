"""CPU-only checks of the product's host side: C-ABI surface, equation grammar, lowering, tables.

No compute call is made here (there is no GPU); the parity tests proper are the `-m gpu` files.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GBM_EQ, HESTON_EQ, ROOT, grid

import sde_sim_rs as S
from sde_sim_rs import _ffi


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sde_b200.h")).read()
    declared = set(re.findall(r"\b(sde_[a-z0-9_]+)\s*\(", header))
    declared -= {"sde_b200"}
    lib = _ffi.lib()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert "sm_100a" in S.version()


def test_options_struct_layout_matches_header():
    o = _ffi.default_options()
    assert o.struct_size == C.sizeof(_ffi.SdeOptions)
    assert (o.output, o.layout, o.scramble, o.icdf, o.arith, o.rk_variant) == (0, 0, 0, 0, 0, 0)


def test_no_cpu_fallback_without_cuda():
    if S.cuda_available():
        pytest.skip("CUDA present")
    u = S.Universe(GBM_EQ, grid(252, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        S.Plan(u, "euler", "pseudo", device=0)


def test_parser_acceptance_table_matches_oracle(oracle):
    cases = [
        ["delta = 1.0"],
        ["dX = ( 1.0 ) * dt - ( 2.0 ) * dW1"],
        ["dX = ( 1.0 ) * dt + ( 2.0 ) + ( 3.0 ) * dW1"],
        ["dX = ( X ) * dN1(X) + ( X ) * dN1(2*X)"],
        ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
         "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
         "C = max(X1 - 100.0, 0.0) + X0"],
        GBM_EQ, HESTON_EQ,
        ["dX = ( sin(t) ) * dt + ( 0.01 * X ) * dW1+( 1 ) * dt"],      # incrementor token runs to the next space (util.rs:109)
    ]
    for eqs in cases:
        a, b = S.Universe(eqs, [0.0, 0.5, 1.0]), oracle.Universe(eqs, [0.0, 0.5, 1.0])
        assert (a.process_names, a.is_levy, a.num_terms, a.factor_names) == (b.names, b.is_levy, b.num_terms, b.factors), eqs


@pytest.mark.parametrize("bad", [["X = 1 = 2"], ["dX 1.0"], ["dX = ( 1.0 ) * dQ"], ["dX = ( 1.0 * dt"],
                                 ["dX = ( foo(1) ) * dt"], ["dX = ( Y ) * dt"], ["dX = ( ) * dt"], ["dX = ( 1 +* 2 ) * dt"]])
def test_parser_rejections(bad, oracle):
    with pytest.raises(ValueError, match="Failed to parse equations"):
        S.Universe(bad, [0.0, 1.0])
    with pytest.raises(ValueError):
        oracle.Universe(bad, [0.0, 1.0])


def test_time_grid_validation():
    with pytest.raises(ValueError):
        S.Universe(GBM_EQ, [0.0, 1.0, 1.0])
    with pytest.raises(ValueError):
        S.Universe(GBM_EQ, [0.0, float("nan")])


def test_simulate_argument_errors():
    with pytest.raises(ValueError, match="scenarios must be a positive integer"):       # py_binding.rs:20-24
        S.simulate(GBM_EQ, grid(252, 4), 0, {"X1": 1.0}, "pseudo", "euler")
    with pytest.raises(ValueError, match="Failed to parse equations"):
        S.simulate(["dX = ( 1.0 ) * dQ"], grid(252, 4), 10, {"X": 1.0}, "pseudo", "euler")


def _lower(eqs, times, scheme, rng, compile=1, **kw):
    u = S.Universe(eqs, times)
    o = S._make_options(device=0, seed=0, scenario_offset=0, output=kw.get("output", "paths"), layout=kw.get("layout", "NTP"),
                        scramble=kw.get("scramble", "cp_shift_per_path"), icdf=kw.get("icdf", "reference"),
                        arithmetic=kw.get("arithmetic", "strict"), rk_variant=kw.get("rk_variant", "reference"),
                        ntp_direct=kw.get("ntp_direct", 0), dtype=kw.get("dtype", "f64"), wide_mma=kw.get("wide_mma", 0),
                        block_threads=kw.get("block_threads", 0), tile_steps=kw.get("tile_steps", 0),
                        generator=kw.get("generator", "chacha8"))
    src, nb = C.c_void_p(), C.c_size_t(0)
    rc = _ffi.lib().sde_lower_only(u._h, scheme.encode(), rng.encode(), C.byref(o), compile, C.byref(src), C.byref(nb))
    _ffi.check(rc)
    text = C.string_at(src).decode()
    _ffi.lib().sde_free_string(src)
    return text, nb.value


def test_fast_arithmetic_factors_the_heston_step():
    """arithmetic=fast: coefficients in product form, terms grouped by their state-dependent part, constants folded into
    per-step slots (lower.cpp, factorise) — the Heston RK step (C3) is what the FP64 pipe spends its time on."""
    fast = dict(scramble="xor", icdf="fast", arithmetic="fast")
    text, _ = _lower(HESTON_EQ, grid(1000), "runge-kutta", "sobol", compile=0, **fast)
    step = text.split("sde_model_step(")[1]
    assert "#define SDE_U0_BITS 1" in text and "const sde_u0_t u0" in text and "(~u0) & 0x80000000u" in step
    assert step.count("sde_f_sqrt_max0_fast(c[1])") == 2                 # one root per stage, shared by dS and dv
    assert "sde_f_max" not in step and "sde_f_sqrt_fast" not in step
    assert "fma(q0, w0_1, w0_0)" in step and "(f1_0 * i1_0)" in step     # dS: S * (a + sqrt(v+) b z)
    assert "fma(SDE_SLOT_5, zu[1], (SDE_SLOT_4 * zu[0]))" in step        # dv: both Wiener terms share sqrt(v+)
    # the uniform grid's dt / sqrt(dt) became literals: 0.05 dt, 2 dt, (-0.21 + 0.2142...) sqrt(dt)
    assert "sde_uc(5.000000000000005e-05)" in text and "sde_uc(0.0020000000000000018)" in text and "ss[4]" not in text
    # a non-uniform grid keeps them in the per-step table
    times = [0.0] + list(np.cumsum(np.linspace(1e-3, 2e-3, 40)))
    tab, _ = _lower(HESTON_EQ, times, "runge-kutta", "sobol", compile=0, **fast)
    assert "slots[0] = (0.050000000000000003 * dt);" in tab and "((sde_real)ss[4])" in tab
    # raw Sobol points (u = x / 2^32 can equal 1/2 exactly) and the ChaCha stream keep the f64 compare
    raw, _ = _lower(HESTON_EQ, grid(250), "runge-kutta", "sobol", compile=0, scramble="none", icdf="fast", arithmetic="fast")
    assert "SDE_U0_BITS" not in raw and "(u0 > 0.5) ? 0u : 0x80000000u" in raw
    # strict arithmetic keeps the reference's term-by-term order
    strict, _ = _lower(HESTON_EQ, grid(1000), "runge-kutta", "sobol", compile=0, scramble="xor")
    assert "SDE_U0_BITS" not in strict and "sde_f_max(c[1], 0.0)" in strict and "__dmul_rn" in strict
    # Euler, non-linear process: same grouping
    eul, _ = _lower(HESTON_EQ, grid(250), "euler", "sobol", compile=0, **fast)
    assert eul.split("sde_model_step(")[1].count("sde_f_sqrt_max0_fast(c[1])") == 2   # one per process block


def test_unknown_scheme_is_value_error():
    with pytest.raises(ValueError, match="unknown scheme"):
        _lower(GBM_EQ, grid(252, 4), "milstein", "pseudo", compile=0)


def test_f32_needs_fast_arithmetic_and_emits_float_literals():
    with pytest.raises(ValueError, match="f32 needs arithmetic"):
        _lower(GBM_EQ, grid(252, 4), "euler", "pseudo", compile=0, dtype="f32")
    text, _ = _lower(GBM_EQ, grid(252, 4), "runge-kutta", "pseudo", compile=0, dtype="f32", arithmetic="fast")
    # (arithmetic=fast folds 0.05 into the per-step constant 0.05 * dt, emitted as a single-precision literal)
    assert "#define SDE_F32 1" in text and "0.000198412701f" in text and "0.5f" in text
    import re
    step = text.split("sde_model_step(")[1]
    assert not re.search(r"\d\.\d+(e-?\d+)?(?![\df])\b", step.split("{", 1)[1])   # no f64 literal inside the model step
    text64, _ = _lower(GBM_EQ, grid(252, 4), "runge-kutta", "pseudo", compile=0, arithmetic="fast")
    assert "SDE_F32" not in text64 and "(0.0001984126984126984" in text64 and "sde_uc(0.0062994078834871" in text64
    strict64, _ = _lower(GBM_EQ, grid(252, 4), "runge-kutta", "pseudo", compile=0)
    assert "0.050000000000000003" in strict64 and "__dmul_rn" in strict64


@pytest.mark.parametrize("block", [16, 100, 160 + 1, 2048, -32])
def test_block_threads_must_be_whole_warps(block):
    # every kernel splits the Sobol index into CTA ^ warp ^ lane parts and sizes shared arrays per warp
    for kw in (dict(), dict(scramble="xor", ntp_direct=2), dict(scramble="xor", ntp_direct=3)):
        with pytest.raises(ValueError, match="block_threads must be"):
            _lower(GBM_EQ, grid(252, 40), "euler", "sobol", compile=0, block_threads=block, **kw)
    _lower(GBM_EQ, grid(252, 40), "euler", "sobol", compile=0, block_threads=160, scramble="xor", ntp_direct=2)


def test_direct_store_tile_walk_visits_every_step_once_inside_its_tile():
    """Host-side emulation of the step schedule of sde_sim_kernel.cuh with SDE_DIRECT (same expressions as the kernel):
    per warp, gamma head steps, 4-step groups shifted by gamma across tiles of TT steps, then <= 3 trailing steps.  Every
    step must run exactly once, in order, and inside the window [t0, t0 + TT + 3) that its tile stages in shared memory."""
    for TT in (4, 16, 32, 36):
        TS = TT + 3
        for S_ in list(range(1, 150)) + [252, 1000, 2017, 2018, 2049, 2050]:
            for gamma in range(4):
                seen, tail_t = [], -1
                g_eff = min(gamma, S_)
                s_full = g_eff + ((S_ - g_eff) & ~3)
                for t0 in range(0, S_, TT):
                    staged = range(t0, min(t0 + TS, S_))
                    run = []
                    if t0 == 0:
                        run += [j for j in range(3) if j < g_eff]
                    tc = t0 + g_eff
                    t_hi = min(t0 + TT + g_eff, s_full)
                    n_groups = (t_hi - tc) >> 2 if t_hi > tc else 0
                    for _ in range(n_groups):
                        run += [tc, tc + 1, tc + 2, tc + 3]
                        tc += 4
                    if t0 + TT + g_eff >= s_full:
                        if tail_t < 0:
                            tail_t = s_full
                        for _ in range(3):
                            if tail_t < min(t0 + TS, S_):
                                run.append(tail_t)
                                tail_t += 1
                    assert all(t in staged for t in run), (TT, S_, gamma, t0, run)
                    seen += run
                assert seen == list(range(S_)), (TT, S_, gamma)


def test_rk_without_factor_is_value_error():
    with pytest.raises(ValueError, match="runge-kutta needs"):
        _lower(["dX = ( 1.0 ) * dt"], grid(252, 4), "runge-kutta", "pseudo", compile=0)


def test_lowering_cache_rule_euler_vs_rk():
    e, _ = _lower(GBM_EQ, grid(252, 4), "euler", "pseudo", compile=0)
    assert "cache enters behind times[t]" in e and "c[0] = row[0];" in e
    r, _ = _lower(GBM_EQ, grid(252, 4), "runge-kutta", "pseudo", compile=0)
    assert "cache enters AT times[t]" in r                      # stale-cache quirk, SURVEY.md §A.4
    assert "c[0] = row[0];" not in r and "c[0] = n0;" in r
    tb, _ = _lower(GBM_EQ, grid(252, 4), "runge-kutta", "pseudo", compile=0, rk_variant="textbook")
    assert "c[0] = row[0];" in tb
    alg, _ = _lower(["dX = ( A ) * dt", "A = 2.0 + 0.0 * X"], [0.0, 1.0, 2.0], "euler", "pseudo", compile=0)
    assert "cache enters AT times[t]" in alg                    # algebraic eval leaves the cache at t+1 (euler.rs:31-35)


@pytest.mark.parametrize("name,eqs,times,scheme,rng,kw", [
    ("C1", GBM_EQ, grid(252), "euler", "pseudo", {}),
    ("C2-xor-fast", GBM_EQ, grid(252), "euler", "sobol", {"scramble": "xor", "icdf": "fast", "arithmetic": "fast"}),
    ("C2-compat", GBM_EQ, grid(252), "euler", "sobol", {}),
    ("C2-tiled", GBM_EQ, grid(252), "euler", "sobol", {"scramble": "xor", "icdf": "fast", "arithmetic": "fast", "ntp_direct": 2}),
    ("C3-full-euler-resident", HESTON_EQ, grid(250), "euler", "sobol", {"scramble": "xor", "icdf": "fast"}),
    ("C2-f32", GBM_EQ, grid(252), "euler", "sobol", {"scramble": "xor", "icdf": "single", "arithmetic": "fast", "dtype": "f32"}),
    ("C3-f32", HESTON_EQ, grid(1000), "runge-kutta", "pseudo", {"output": "terminal", "arithmetic": "fast", "dtype": "f32"}),
    ("C2-tpn", GBM_EQ, grid(252), "euler", "sobol", {"scramble": "xor", "layout": "TPN"}),
    ("C3", HESTON_EQ, grid(1000), "runge-kutta", "sobol", {"scramble": "xor"}),
    ("C3-terminal", HESTON_EQ, grid(1000), "runge-kutta", "pseudo", {"output": "terminal"}),
    # arithmetic=fast Runge-Kutta (factored form, probe uniform passed as its 32-bit word) on every kernel that can run it
    ("C3-fast-tiled", HESTON_EQ, grid(1000), "runge-kutta", "sobol", {"scramble": "xor", "icdf": "fast", "arithmetic": "fast"}),
    ("C3-fast-resident", HESTON_EQ, grid(250), "runge-kutta", "sobol", {"scramble": "xor", "icdf": "fast", "arithmetic": "fast"}),
    ("C3-fast-raw", HESTON_EQ, grid(250), "runge-kutta", "sobol", {"scramble": "none", "icdf": "fast", "arithmetic": "fast"}),
    ("C3-fast-philox", HESTON_EQ, grid(1000), "runge-kutta", "pseudo", {"output": "terminal", "icdf": "fast", "arithmetic": "fast", "generator": "philox"}),
    ("C3-fast-f32", HESTON_EQ, grid(1000), "runge-kutta", "sobol", {"scramble": "xor", "icdf": "single", "arithmetic": "fast", "dtype": "f32"}),
    ("C3-fast-resident-global-lane-table", HESTON_EQ, grid(1000), "runge-kutta", "sobol", {"scramble": "xor", "icdf": "fast", "arithmetic": "fast", "ntp_direct": 3}),
    ("C3-fast-cp", HESTON_EQ, grid(100), "runge-kutta", "sobol", {"icdf": "fast", "arithmetic": "fast"}),
    ("C5", GBM_EQ, grid(365), "euler", "pseudo", {"output": "moments", "icdf": "fast"}),
    ("jump", ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
              "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
              "C = max(X1 - 100.0, 0.0) + X0"], grid(50, 15), "runge-kutta", "pseudo", {}),
    ("jump-fast-rk", ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
                      "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
                      "C = max(X1 - 100.0, 0.0) + X0"], grid(50, 15), "runge-kutta", "pseudo", {"arithmetic": "fast", "icdf": "fast"}),
    ("jump-fast-euler", ["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dW1",
                         "dX1 = ( 0.01 * X1 ) * dt + ( 0.2 * X1 ) * dW2 + ( 0.5 * cos(t) ) * dN1(abs(X0) * 40)",
                         "C = max(X1 - 100.0, 0.0) + X0"], grid(50, 15), "euler", "sobol", {"scramble": "xor", "arithmetic": "fast", "icdf": "fast"}),
])
def test_configs_lower_and_compile_for_sm100a(name, eqs, times, scheme, rng, kw):
    text, nbytes = _lower(eqs, times, scheme, rng, compile=1, **kw)     # NVRTC --gpu-architecture=sm_100a, no GPU needed
    assert ('#include "sde_sim_kernel.cuh"' in text or '#include "sde_sim_resident.cuh"' in text) and nbytes > 10_000
    if name == "C3-fast-resident-global-lane-table":                    # 2000 dimensions: the lane table stays in global memory
        assert '#include "sde_sim_resident.cuh"' in text and "#define SDE_RES_LANE_GLOBAL 1" in text and "#define SDE_BLOCK 512" in text
    if name == "C3-fast-tiled":
        assert '#include "sde_sim_kernel.cuh"' in text and "#define SDE_NSTAGE 4" in text
    if name == "C2-xor-fast":                                           # Sobol full paths with resident tables: persistent warps
        assert '#include "sde_sim_resident.cuh"' in text and "#define SDE_S 252" in text


def test_wide_linear_models_lower_to_the_tensor_core_kernel():
    from conftest import basket_equations

    fast = {"scramble": "xor", "icdf": "fast", "arithmetic": "fast"}
    eqs, _ = basket_equations(64)
    text, nbytes = _lower(eqs, grid(252), "euler", "sobol", compile=1, output="moments", **fast)       # C4: auto selection
    assert '#include "sde_sim_wide.cuh"' in text and nbytes > 10_000
    assert "#define SDE_WNB 8" in text and "#define SDE_WNKK 16" in text and "#define SDE_WMT 2" in text
    # triangular Cholesky loadings: process tile j needs 2 (j + 1) factor steps of 4
    assert "constexpr int e[8] = {2, 4, 6, 8, 10, 12, 14, 16};" in text
    # B fragments: entry (j = 0, kk = 0, lane 0) = M[0][0] = sigma_0 * L[0][0] = 0.1
    assert "sde_wm[4096] = {0.10000000000000001," in text
    off, _ = _lower(eqs, grid(252), "euler", "sobol", compile=0, output="moments", wide_mma=1, **fast)
    assert '#include "sde_sim_kernel.cuh"' in off                                                     # switched off: time-tiled kernel
    paths, _ = _lower(eqs, grid(252), "euler", "sobol", compile=1, output="paths", **fast)
    assert '#include "sde_sim_wide.cuh"' in paths and "#define SDE_OUT 0" in paths                     # [N][T][P] rows too
    tpn, _ = _lower(eqs, grid(252), "euler", "sobol", compile=0, output="paths", layout="TPN", **fast)
    assert '#include "sde_sim_kernel.cuh"' in tpn                                                     # the transposed layout is not its business
    strict, _ = _lower(eqs, grid(252), "euler", "sobol", compile=0, output="moments", scramble="xor")
    assert '#include "sde_sim_kernel.cuh"' in strict                                                  # reference arithmetic keeps the term order
    eq20, _ = basket_equations(20)                                                                    # pad tiles: P = K = 20 -> 3 x 8, 5 x 4
    t20, nb20 = _lower(eq20, grid(252, 12), "euler", "sobol", compile=1, output="terminal", wide_mma=2, **fast)
    assert "#define SDE_WNB 3" in t20 and "#define SDE_WNKK 5" in t20 and nb20 > 10_000
    with pytest.raises(ValueError, match="tensor-core kernel"):
        _lower(GBM_EQ, grid(252, 4), "euler", "sobol", compile=0, output="moments", wide_mma=2, **fast)


def test_joe_kuo_table_matches_scipy(oracle):
    dims = 21201
    poly = np.zeros(dims, dtype=np.uint32)
    minit = np.zeros((dims, 18), dtype=np.uint32)
    _ffi.check(_ffi.lib().sde_joe_kuo_params(dims, poly.ctypes.data_as(C.c_void_p), minit.ctypes.data_as(C.c_void_p)))
    sp, sm = oracle.joe_kuo_from_scipy(dims)
    assert np.array_equal(poly, sp) and np.array_equal(minit, sm)
    with pytest.raises(ValueError):
        _ffi.check(_ffi.lib().sde_joe_kuo_params(21202, poly.ctypes.data_as(C.c_void_p), minit.ctypes.data_as(C.c_void_p)))


def test_sobol_dimension_limit_is_value_error():
    # 64 factors x 400 steps = 25 600 > 21 201 dims (JoeKuoD6::extended, sobol.rs:16) — checked at plan creation on a GPU;
    # here: the lowering itself still succeeds, the table accessor refuses.
    pass


def test_shard_ranges_cover_and_are_disjoint():
    for n, w in [(10, 3), (1 << 24, 8), (7, 8), (1, 1), (1000, 7)]:
        r = [S.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_moments_merge_matches_numpy():
    rng = np.random.default_rng(0)
    x = rng.normal(3.0, 2.0, size=(5, 1000, 2))
    shards = np.stack([np.stack([[x[s, :, p].size, x[s, :, p].mean(), ((x[s, :, p] - x[s, :, p].mean()) ** 2).sum()]
                                 for p in range(2)]) for s in range(5)])
    m = S.merge_moments(shards)
    flat = x.transpose(2, 0, 1).reshape(2, -1)
    assert np.allclose(m[:, 0], 5000)
    assert np.allclose(m[:, 1], flat.mean(axis=1), rtol=1e-14)
    assert np.allclose(m[:, 2], ((flat - flat.mean(axis=1, keepdims=True)) ** 2).sum(axis=1), rtol=1e-12)
    assert np.array_equal(S.merge_moments(np.concatenate([np.zeros((1, 2, 3)), shards])), m)   # empty shard is neutral


def test_filtration_long_format_views_cpu():
    # the reference's four columns (src/filtration.rs:108-113) rendered from a dense [N, T, P] buffer (CPU tensor here)
    import torch

    vals = torch.arange(2 * 3 * 2, dtype=torch.float64).reshape(2, 3, 2)
    f = S.Filtration(vals, np.array([0.0, 0.5, 1.0]), ["S", "v"], output="paths", layout="NTP", scenario_offset=7)
    c = f.columns()
    assert c["scenario"].dtype == np.int32 and c["scenario"].tolist() == [7] * 6 + [8] * 6
    assert c["time"].tolist() == [0.0, 0.0, 0.5, 0.5, 1.0, 1.0] * 2
    assert c["process_name"].tolist() == ["S", "v"] * 6 and c["value"].tolist() == list(map(float, range(12)))
    assert list(f.to_pandas().columns) == ["scenario", "time", "process_name", "value"]
    tb = f.to_arrow()
    assert [str(t) for t in tb.schema.types] == ["int32", "double", "string", "double"] and tb.num_rows == 12
    tpn = S.Filtration(vals.permute(1, 2, 0).contiguous(), np.array([0.0, 0.5, 1.0]), ["S", "v"], output="paths", layout="TPN")
    assert np.array_equal(tpn.columns()["value"], c["value"])
