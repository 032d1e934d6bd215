"""ctypes wrapper around oracle/libsde_oracle.so — TEST INFRASTRUCTURE ONLY.

The CPU oracle restates the reference hot path (see the header of sde_oracle.cpp;
"parity unpinned": no Rust toolchain / golden vectors exist to pin it to the crate).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under sde-sim-rs_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsde_oracle.so")

RNG_PSEUDO, RNG_SOBOL_CP_SHIFT, RNG_SOBOL_XOR, RNG_SOBOL_RAW, RNG_INJECT, RNG_PHILOX = 0, 1, 2, 3, 4, 5
SCHEME_EULER, SCHEME_RK, SCHEME_RK_TEXTBOOK = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no reference sources)."""
    src = os.path.join(_HERE, "sde_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u64, dbl, i32, vp, sz = C.c_uint64, C.c_double, C.c_int, C.c_void_p, C.c_size_t
        L.orc_sobol_direction_numbers.argtypes = [vp, vp, i32, i32, vp]
        L.orc_sobol_points_sequential.argtypes = [vp, i32, u64, u64, vp]
        L.orc_sobol_points_direct.argtypes = [vp, i32, u64, u64, vp]
        L.orc_chacha_block.argtypes = [vp, u64, u64, i32, vp]
        L.orc_seed_from_u64.argtypes = [u64, vp]
        L.orc_chacha8_u64_stream.argtypes = [u64, sz, vp]
        L.orc_chacha8_f64_stream.argtypes = [u64, sz, vp]
        L.orc_icdf_normal.argtypes = [dbl]
        L.orc_icdf_normal.restype = dbl
        L.orc_icdf_normal_array.argtypes = [vp, sz, vp]
        L.orc_icdf_poisson.argtypes = [dbl, dbl]
        L.orc_icdf_poisson.restype = u64
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_philox_uniform.argtypes = [u64, u64, u64]
        L.orc_philox_uniform.restype = dbl
        L.orc_xor_uniform.argtypes = [u64, u64]
        L.orc_xor_uniform.restype = dbl
        L.orc_last_error.restype = C.c_char_p
        L.orc_universe_parse.argtypes = [C.POINTER(C.c_char_p), i32, vp, i32]
        L.orc_universe_parse.restype = vp
        L.orc_universe_free.argtypes = [vp]
        for f in ("orc_universe_num_processes", "orc_universe_num_factors"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = i32
        for f in ("orc_universe_process_is_levy", "orc_universe_num_terms"):
            getattr(L, f).argtypes = [vp, i32]
            getattr(L, f).restype = i32
        for f in ("orc_universe_process_name", "orc_universe_factor_name"):
            getattr(L, f).argtypes = [vp, i32]
            getattr(L, f).restype = C.c_char_p
        L.orc_expr_eval.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), vp, i32, dbl, C.POINTER(dbl)]
        L.orc_expr_eval.restype = i32
        L.orc_simulate.argtypes = [vp, C.POINTER(C.c_char_p), vp, i32, u64, i32, i32, u64, u64, vp, vp, i32, vp]
        L.orc_simulate.restype = i32
        L.orc_num_threads.restype = i32
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------- Joe–Kuo parameters
def joe_kuo_from_scipy(dims: int):
    """(poly[dims] u32, minit[dims][18] u32) straight from scipy's copy of new-joe-kuo-6.21201.

    scipy row i is Sobol dimension i+1 (row 0 = van der Corput).  This is the independent
    source the oracle is pinned to; the product ships its own derived copy under
    sde-sim-rs_b200/data/ which tests compare against this.
    """
    import scipy

    z = np.load(os.path.join(os.path.dirname(scipy.__file__), "stats", "_sobol_direction_numbers.npz"))
    poly = np.ascontiguousarray(z["poly"][:dims].astype(np.uint32))
    minit = np.ascontiguousarray(z["vinit"][:dims].astype(np.uint32))
    return poly, minit


def sobol_direction_numbers(dims: int) -> np.ndarray:
    poly, minit = joe_kuo_from_scipy(dims)
    V = np.zeros((dims, 64), dtype=np.uint64)
    lib().orc_sobol_direction_numbers(_ptr(poly), _ptr(minit), minit.shape[1], dims, _ptr(V))
    return V


def sobol_points(V: np.ndarray, first: int, count: int, sequential: bool = False) -> np.ndarray:
    dims = V.shape[0]
    out = np.zeros((count, dims), dtype=np.uint64)
    fn = lib().orc_sobol_points_sequential if sequential else lib().orc_sobol_points_direct
    fn(_ptr(V), dims, first, count, _ptr(out))
    return out


# ---------------------------------------------------------------- ChaCha
def chacha_block(key8, counter: int, rounds: int, stream: int = 0) -> np.ndarray:
    key = np.ascontiguousarray(np.asarray(key8, dtype=np.uint32))
    out = np.zeros(16, dtype=np.uint32)
    lib().orc_chacha_block(_ptr(key), counter, stream, rounds, _ptr(out))
    return out


def seed_from_u64(seed: int) -> np.ndarray:
    key = np.zeros(8, dtype=np.uint32)
    lib().orc_seed_from_u64(seed & (2**64 - 1), _ptr(key))
    return key


def chacha8_u64(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_chacha8_u64_stream(seed & (2**64 - 1), n, _ptr(out))
    return out


def chacha8_f64(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float64)
    lib().orc_chacha8_f64_stream(seed & (2**64 - 1), n, _ptr(out))
    return out


# ---------------------------------------------------------------- inverse CDFs
def icdf_normal(p):
    p = np.ascontiguousarray(np.asarray(p, dtype=np.float64))
    out = np.empty_like(p)
    lib().orc_icdf_normal_array(_ptr(p), p.size, _ptr(out))
    return out


def icdf_poisson(u: float, lam: float) -> int:
    return int(lib().orc_icdf_poisson(u, lam))


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.ascontiguousarray(np.asarray(ctr, dtype=np.uint32))
    k = np.ascontiguousarray(np.asarray(key, dtype=np.uint32))
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def philox_uniform(seed: int, scenario: int, draw: int) -> float:
    return float(lib().orc_philox_uniform(seed & (2**64 - 1), scenario, draw))


def xor_uniform(x: int, mask: int) -> float:
    return float(lib().orc_xor_uniform(x, mask))


# ---------------------------------------------------------------- expressions / model
def _cstrs(strs):
    arr = (C.c_char_p * len(strs))(*[s.encode() for s in strs])
    return arr


def expr_eval(src: str, values: dict | None = None, t: float = 0.0) -> float:
    values = values or {}
    names = list(values)
    vals = np.asarray([values[k] for k in names], dtype=np.float64)
    out = C.c_double()
    rc = lib().orc_expr_eval(src.encode(), _cstrs(names), _ptr(vals), len(names), t, C.byref(out))
    if rc:
        raise ValueError(lib().orc_last_error().decode())
    return out.value


class Universe:
    """proc::util::parse_equations restated (src/proc/util.rs:52-66)."""

    def __init__(self, equations, time_steps):
        self.times = np.ascontiguousarray(np.asarray(time_steps, dtype=np.float64))
        self._h = lib().orc_universe_parse(_cstrs(list(equations)), len(equations), _ptr(self.times), self.times.size)
        if not self._h:
            raise ValueError("Failed to parse equations: " + lib().orc_last_error().decode())
        L = lib()
        self.P = L.orc_universe_num_processes(self._h)
        self.K = L.orc_universe_num_factors(self._h)
        self.names = [L.orc_universe_process_name(self._h, i).decode() for i in range(self.P)]
        self.is_levy = [bool(L.orc_universe_process_is_levy(self._h, i)) for i in range(self.P)]
        self.num_terms = [L.orc_universe_num_terms(self._h, i) for i in range(self.P)]
        self.factors = [L.orc_universe_factor_name(self._h, i).decode() for i in range(self.K)]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.orc_universe_free(h)


def simulate(universe: Universe, initial_values: dict, scenarios: int, scheme: str = "euler",
             rng_method: str = "pseudo", *, seed: int = 0, scramble: str = "cp_shift_per_path",
             scenario_offset: int = 0, inject: np.ndarray | None = None, nthreads: int = 0,
             textbook_rk: bool = False, generator: str = "chacha8") -> np.ndarray:
    """sim::simulate restated (src/sim/mod.rs:20-92) -> dense [N, T, P] f64."""
    T, P, K = universe.times.size, universe.P, universe.K
    S = T - 1
    if scheme == "euler":
        sch = SCHEME_EULER
    elif scheme == "runge-kutta":
        sch = SCHEME_RK_TEXTBOOK if textbook_rk else SCHEME_RK
    else:
        raise ValueError("unknown scheme (reference: unimplemented!() panic, src/sim/mod.rs:82)")
    V = None
    if inject is not None:
        mode = RNG_INJECT
        inject = np.ascontiguousarray(inject, dtype=np.float64)
        assert inject.shape == (scenarios, S, K + 1), (inject.shape, (scenarios, S, K + 1))
    elif rng_method == "sobol":
        mode = {"cp_shift_per_path": RNG_SOBOL_CP_SHIFT, "xor": RNG_SOBOL_XOR, "none": RNG_SOBOL_RAW}[scramble]
        if S * K > 0:
            V = sobol_direction_numbers(S * K)
    elif generator == "philox":
        mode = RNG_PHILOX                                   # not in the reference: the engine's counter-based MC tier
    else:
        mode = RNG_PSEUDO                                   # any other string -> pseudo (src/sim/mod.rs:65)
    names = list(initial_values)
    vals = np.asarray([initial_values[k] for k in names], dtype=np.float64)
    out = np.zeros((scenarios, T, P), dtype=np.float64)
    rc = lib().orc_simulate(universe._h, _cstrs(names), _ptr(vals), len(names), scenarios, sch, mode,
                            seed & (2**64 - 1), scenario_offset, _ptr(V), _ptr(inject), nthreads, _ptr(out))
    if rc:
        raise RuntimeError("Simulation failed: " + lib().orc_last_error().decode())
    return out


def uniforms(universe: Universe, scenarios: int, rng_method: str, *, seed: int = 0,
             scramble: str = "cp_shift_per_path", scenario_offset: int = 0, generator: str = "chacha8") -> np.ndarray:
    """The u[s][t][k] stream a run would consume (for building injected-normal inputs)."""
    S, K = universe.times.size - 1, universe.K
    out = np.zeros((scenarios, S, K), dtype=np.float64)
    if rng_method != "sobol" and generator == "philox":
        for s in range(scenarios):
            out[s] = np.array([philox_uniform(seed, s + scenario_offset, i) for i in range(S * K)]).reshape(S, K)
        return out
    if rng_method != "sobol":
        for s in range(scenarios):
            out[s] = chacha8_f64(s + scenario_offset + seed, S * K).reshape(S, K)
        return out
    V = sobol_direction_numbers(S * K)
    pts = sobol_points(V, scenario_offset + 5, scenarios)
    if scramble == "none":
        return (pts.astype(np.float64) * 2.0**-64).reshape(scenarios, S, K)
    if scramble == "xor":
        masks = chacha8_u64(seed, S * K)
        k = (pts ^ masks[None, :]) >> np.uint64(32)
        return ((k.astype(np.float64) + 0.5) * 2.0**-32).reshape(scenarios, S, K)
    raw = pts.astype(np.float64) * 2.0**-64                 # exact for n < 2^53
    for s in range(scenarios):
        v = raw[s] + chacha8_f64(s + scenario_offset + seed, S * K)
        out[s] = (v - np.trunc(v)).reshape(S, K)
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())
