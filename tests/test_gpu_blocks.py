"""GPU parity of the building blocks (each kernel on its own) against the CPU oracle, through the C-ABI.

Bars: integer/bit work bit-exact; inverse normal CDF within the tolerance stated below.
"""
import ctypes as C

import numpy as np
import pytest

from sde_sim_rs import _ffi

pytestmark = pytest.mark.gpu


def _sobol_dev(dims, first, count):
    out = np.zeros((count, dims), dtype=np.uint64)
    _ffi.check(_ffi.lib().sde_sobol_points(0, dims, first, count, out.ctypes.data_as(C.c_void_p)))
    return out


@pytest.mark.parametrize("dims,first,count", [
    (252, 5, 4096),                 # C2: first point the reference uses (skip(5), sobol.rs:17)
    (252, 0, 300),
    (252, (1 << 24) + 5 - 100, 200),   # across the 2^24 boundary = shard boundary of an 8-GPU C2-size run
    (252, (1 << 32) - 1000, 1000),  # top of the supported index range
    (2000, 5, 512),                 # C3
    (16128, 5, 64),                 # C4
    (21201, 1234567, 32),           # every dimension of the Joe–Kuo table
    (1, 0, 1), (3, 255, 2), (7, 256, 257),
])
def test_sobol_points_bit_exact(oracle, dims, first, count):
    V = oracle.sobol_direction_numbers(dims)
    ref = oracle.sobol_points(V, first, count)
    got = _sobol_dev(dims, first, count)
    assert np.array_equal(got, ref)


def test_sobol_points_match_scipy_directly():
    from scipy.stats import qmc

    got = _sobol_dev(64, 0, 128).astype(np.float64) * 2.0**-64
    assert np.array_equal(got, qmc.Sobol(d=64, scramble=False, bits=64).random(128))


@pytest.mark.parametrize("dims,seed,first,count", [(252, 42, 0, 300), (504, 7, 1000, 64), (2000, 2**63 + 5, 4091, 16), (3, 0, 2**24 - 5, 10)])
def test_cp_shift_per_path_uniforms_bit_exact(oracle, dims, seed, first, count):
    # the reference's Sobol mode (SobolRng::new + RandomShiftScrambler, src/rng/sobol.rs:35-53,62-79): u = fract(x_n[d] + ChaCha8(seed + s).f64[d])
    out = np.zeros((count, dims), dtype=np.float64)
    _ffi.check(_ffi.lib().sde_sobol_cp_shift_uniforms(0, dims, seed, first, count, out.ctypes.data_as(C.c_void_p)))
    U = oracle.Universe(["dA = ( 1.0 ) * dW1"], np.arange(dims + 1, dtype=np.float64))
    ref = oracle.uniforms(U, count, "sobol", seed=seed, scramble="cp_shift_per_path", scenario_offset=first).reshape(count, dims)
    assert np.array_equal(out, ref)
    assert out.min() >= 0.0 and out.max() < 1.0


@pytest.mark.parametrize("seed,n", [(0, 64), (42, 1000), (2**64 - 1, 17), (123456789, 8)])
def test_chacha8_stream_bit_exact(oracle, seed, n):
    out = np.zeros(n, dtype=np.uint64)
    _ffi.check(_ffi.lib().sde_chacha8_u64(0, seed, n, out.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(out, oracle.chacha8_u64(seed, n))


def _icdf_dev(p, mode):
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(p)
    _ffi.check(_ffi.lib().sde_icdf_normal(0, mode, p.ctypes.data_as(C.c_void_p), p.size, out.ctypes.data_as(C.c_void_p)))
    return out


def _icdf_inputs():
    rng = np.random.default_rng(1)
    p = np.concatenate([
        rng.random(200_000),
        (rng.integers(1, 2**53, size=50_000, dtype=np.uint64) >> np.uint64(11)).astype(np.float64) * 2.0**-53 + 2.0**-53,
        2.0 ** -np.arange(1, 54, dtype=np.float64),                     # down to the smallest 53-bit uniform
        1.0 - 2.0 ** -np.arange(1, 54, dtype=np.float64),
        np.linspace(0.499, 0.501, 2001),
        [0.5, 0.975, 0.025, 0.875, 0.375, 1e-9, 0.7090754154265618],
    ])
    return p[(p > 0) & (p < 1)]


def test_icdf_reference_mode_tolerance(oracle):
    # Stated tolerance (device REFERENCE vs oracle): <= 4 ulp relative away from p ~ 0.5 and <= 1e-15 absolute
    # near it (the result there is a ~1e-7 cancellation residue).  Only CUDA log vs glibc log can differ.
    p = _icdf_inputs()
    ref, got = oracle.icdf_normal(p), _icdf_dev(p, 0)
    err = np.abs(got - ref)
    tol = np.maximum(1e-15, 4 * np.spacing(np.abs(ref)))
    assert np.all(err <= tol), (err.max(), p[np.argmax(err - tol)])


def test_icdf_fast_mode_tolerance(oracle):
    # Stated tolerance (device FAST vs oracle): |dz| <= 5e-13 absolute over p in [2^-53, 1 - 2^-53].
    p = _icdf_inputs()
    ref, got = oracle.icdf_normal(p), _icdf_dev(p, 1)
    err = np.abs(got - ref)
    print("fast icdf max abs err", err.max(), "at p =", p[np.argmax(err)])
    assert err.max() <= 5e-13


def test_icdf_fast_k32_front_end_tolerance(oracle):
    # The digital-shift path feeds 32-bit integers k (p = (k + 1/2) 2^-32) to the integer front end, and the persistent
    # kernel forms the exponent term arithmetically (sde_icdf_normal_fast_k32s).  Same stated tolerance: 5e-13 absolute.
    # Every leading-one position of min(p, 1-p) (all 32 exponents, both signs), their neighbours, and random k.
    rng = np.random.default_rng(2)
    k = [np.uint64(0), np.uint64(2**32 - 1), np.uint64(2**31), np.uint64(2**31 - 1)]
    for b in range(32):
        for d in (-1, 0, 1):
            for base in (2**b, 2**32 - 1 - 2**b):
                v = base + d
                if 0 <= v < 2**32:
                    k.append(np.uint64(v))
    k = np.concatenate([np.array(k, dtype=np.uint64), rng.integers(0, 2**32, size=400_000, dtype=np.uint64),
                        rng.integers(0, 2**12, size=20_000, dtype=np.uint64)])          # deep left tail
    p = (k.astype(np.float64) + 0.5) * 2.0**-32
    ref = oracle.icdf_normal(p)
    # tiled kernel / persistent kernel / persistent kernel with FP32-unit seeds + quadratic steps (sign-folded entry)
    for mode, name in ((2, "128-entry table"), (5, "1024-entry table"), (6, "1024-entry table, FP32-unit seeds")):
        got = _icdf_dev(p, mode)
        err = np.abs(got - ref)
        print("k32 icdf,", name, "max abs err", err.max(), "at k =", int(k[np.argmax(err)]))
        assert err.max() <= 5e-13
        assert np.array_equal(np.signbit(got), np.signbit(ref))      # x(w) itself is slightly negative next to p = 1/2


def test_icdf_single_precision_tier_tolerance(oracle):
    # icdf="single": the A&S map in FP32.  Stated tolerance vs the reference evaluation: |dz| <= 4e-6 absolute
    # everywhere, <= 1e-6 for |z| <= 3 (both entries: f64 uniforms and the 32-bit integer front end).
    p = _icdf_inputs()
    p = p[(p >= 2.0**-33) & (p <= 1 - 2.0**-33)]                       # the range a 32-bit uniform can reach
    ref, got = oracle.icdf_normal(p), _icdf_dev(p, 3)
    err = np.abs(got - ref)
    print("single icdf (f64 entry) max abs err", err.max(), "central", err[np.abs(ref) <= 3].max())
    assert err.max() <= 4e-6 and err[np.abs(ref) <= 3].max() <= 1e-6
    rng = np.random.default_rng(3)
    k = np.concatenate([rng.integers(0, 2**32, size=300_000, dtype=np.uint64), rng.integers(0, 2**10, size=5_000, dtype=np.uint64),
                        np.array([0, 1, 2**31 - 1, 2**31, 2**32 - 1], dtype=np.uint64)])
    pk = (k.astype(np.float64) + 0.5) * 2.0**-32
    ref, got = oracle.icdf_normal(pk), _icdf_dev(pk, 4)
    err = np.abs(got - ref)
    print("single icdf (k32 entry) max abs err", err.max(), "central", err[np.abs(ref) <= 3].max())
    assert err.max() <= 4e-6 and err[np.abs(ref) <= 3].max() <= 1e-6
    assert np.isnan(_icdf_dev([0.0], 3)[0])


def test_icdf_zero_is_nan_both_modes():
    assert np.isnan(_icdf_dev([0.0], 0)[0]) and np.isnan(_icdf_dev([0.0], 1)[0])    # ln(0) path, increment.rs:165-177


def test_icdf_poisson_exact(oracle):
    rng = np.random.default_rng(2)
    u = np.concatenate([rng.random(20000), [0.1, 0.95, 0.96, 0.999, 0.9999, 0.01, 0.5, 0.99, 1.0, 0.5, 0.5]])
    lam = np.concatenate([rng.random(20000) * 6.0, [0.05] * 5, [3.0] * 3, [500.0, 0.0, -1.0]])
    out = np.empty_like(u)
    _ffi.check(_ffi.lib().sde_icdf_poisson(0, u.ctypes.data_as(C.c_void_p), lam.ctypes.data_as(C.c_void_p), u.size,
                                           out.ctypes.data_as(C.c_void_p)))
    ref = np.array([oracle.icdf_poisson(float(a), float(b)) for a, b in zip(u, lam)], dtype=np.float64)
    # exp() may differ by an ulp between CUDA and glibc: allow a count difference only where u sits within 1e-13 of a CDF step
    bad = np.flatnonzero(out != ref)
    assert bad.size <= 2, bad.size
