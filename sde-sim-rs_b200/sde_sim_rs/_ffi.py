"""ctypes binding of include/sde_b200.h (libsde_b200.so, built in-tree by ../build.py).

This is the Python side of the drop-in boundary: the functions below are what the
reference's pyo3 module (src/py_binding.rs) would bind if its hot path were replaced by
this library.  The library is the only compute path: if it is missing, import fails loudly;
if CUDA is missing, every compute call raises RuntimeError.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsde_b200.so")

SDE_OK, SDE_ERR_VALUE, SDE_ERR_RUNTIME = 0, 1, 2

OUT_PATHS, OUT_TERMINAL, OUT_MOMENTS = 0, 1, 2
LAYOUT_NTP, LAYOUT_TPN = 0, 1
SCRAMBLE_CP_SHIFT_PER_PATH, SCRAMBLE_XOR, SCRAMBLE_NONE = 0, 1, 2
ICDF_REFERENCE, ICDF_FAST, ICDF_SINGLE = 0, 1, 2
DTYPE_F64, DTYPE_F32 = 0, 1
ARITH_STRICT, ARITH_FAST = 0, 1
RK_REFERENCE, RK_TEXTBOOK = 0, 1
GEN_CHACHA8, GEN_PHILOX = 0, 1


class SdeOptions(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("seed", C.c_uint64),
        ("scenario_offset", C.c_uint64),
        ("output", C.c_int32),
        ("layout", C.c_int32),
        ("scramble", C.c_int32),
        ("icdf", C.c_int32),
        ("arith", C.c_int32),
        ("rk_variant", C.c_int32),
        ("stream", C.c_void_p),
        ("inject", C.c_void_p),
        ("tile_steps", C.c_int32),
        ("block_threads", C.c_int32),
        ("min_blocks", C.c_int32),
        ("ntp_direct", C.c_int32),
        ("dtype", C.c_int32),
        ("wide_mma", C.c_int32),
        ("generator", C.c_int32),
    ]


# name -> (restype, argtypes): every symbol include/sde_b200.h declares
_vp, _sz, _u64, _u32, _i32, _dbl = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, C.c_double
_strs = C.POINTER(C.c_char_p)
_popt = C.POINTER(SdeOptions)
_pint = C.POINTER(C.c_int)
SIGNATURES = {
    "sde_universe_parse": (_i32, [_strs, _sz, _vp, _sz, C.POINTER(_vp)]),
    "sde_universe_free": (None, [_vp]),
    "sde_universe_num_processes": (_sz, [_vp]),
    "sde_universe_num_factors": (_sz, [_vp]),
    "sde_universe_num_times": (_sz, [_vp]),
    "sde_universe_process_name": (C.c_char_p, [_vp, _sz]),
    "sde_universe_process_is_levy": (_i32, [_vp, _sz]),
    "sde_universe_process_num_terms": (_sz, [_vp, _sz]),
    "sde_universe_factor_name": (C.c_char_p, [_vp, _sz]),
    "sde_options_default": (None, [_popt]),
    "sde_plan_create": (_i32, [_vp, C.c_char_p, C.c_char_p, _popt, C.POINTER(_vp)]),
    "sde_plan_free": (None, [_vp]),
    "sde_plan_source": (C.c_char_p, [_vp]),
    "sde_plan_is_prelowered": (_i32, [_vp]),
    "sde_plan_output_elems": (_sz, [_vp, _u64]),
    "sde_lower_only": (_i32, [_vp, C.c_char_p, C.c_char_p, _popt, _i32, C.POINTER(_vp), C.POINTER(_sz)]),
    "sde_free_string": (None, [_vp]),
    "sde_plan_run_device": (_i32, [_vp, _strs, _vp, _sz, _u64, _u64, _u64, _vp, _vp, _pint]),
    "sde_plan_run_host": (_i32, [_vp, _strs, _vp, _sz, _u64, _u64, _u64, _vp, _pint]),
    "sde_simulate": (_i32, [_vp, _strs, _vp, _sz, _u64, C.c_char_p, C.c_char_p, _popt, C.POINTER(_vp)]),
    "sde_result_free": (None, [_vp]),
    "sde_result_shape": (None, [_vp, C.POINTER(_u64), C.POINTER(_sz), C.POINTER(_sz)]),
    "sde_result_num_elems": (_sz, [_vp]),
    "sde_result_values_device": (_vp, [_vp]),
    "sde_result_values_host": (_i32, [_vp, _vp, _sz]),
    "sde_result_kernel_ms": (_dbl, [_vp]),
    "sde_result_process_name": (C.c_char_p, [_vp, _sz]),
    "sde_result_times": (_i32, [_vp, _vp, _sz]),
    "sde_result_scenario_offset": (_u64, [_vp]),
    "sde_result_device": (_i32, [_vp]),
    "sde_result_output": (_i32, [_vp]),
    "sde_result_moments": (_i32, [_vp, _vp]),
    "sde_shard_range": (None, [_u64, _sz, _sz, C.POINTER(_u64), C.POINTER(_u64)]),
    "sde_device_plans_create": (_i32, [_vp, C.c_char_p, C.c_char_p, _popt, C.POINTER(C.c_int32), _sz, C.POINTER(_vp)]),
    "sde_device_plans_free": (None, [_vp]),
    "sde_device_plans_count": (_sz, [_vp]),
    "sde_device_plans_device": (_i32, [_vp, _sz]),
    "sde_device_plans_collective": (_i32, [_vp]),
    "sde_plan_run_devices": (_i32, [_vp, _strs, _vp, _sz, _u64, _u64, _u64, C.POINTER(_vp), _pint, C.POINTER(_dbl)]),
    "sde_simulate_devices": (_i32, [_vp, _strs, _vp, _sz, _u64, C.c_char_p, C.c_char_p, _popt, C.POINTER(C.c_int32), _sz, C.POINTER(_vp)]),
    "sde_moments_merge_device": (_i32, [_i32, _vp, _sz, _sz, _vp, _vp]),
    "sde_sobol_points": (_i32, [_i32, _u32, _u64, _u64, _vp]),
    "sde_sobol_cp_shift_uniforms": (_i32, [_i32, _u32, _u64, _u64, _u64, _vp]),
    "sde_joe_kuo_params": (_i32, [_u32, _vp, _vp]),
    "sde_chacha8_u64": (_i32, [_i32, _u64, _sz, _vp]),
    "sde_icdf_normal": (_i32, [_i32, _i32, _vp, _sz, _vp]),
    "sde_icdf_poisson": (_i32, [_i32, _vp, _vp, _sz, _vp]),
    "sde_moments_merge": (_i32, [_vp, _sz, _sz, _vp]),
    "sde_measure_peaks": (_i32, [_i32, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl)]),
    "sde_last_error": (C.c_char_p, []),
    "sde_version": (C.c_char_p, []),
    "sde_cuda_available": (_i32, []),
    "sde_device_count": (_i32, []),
}

_lib = None

# compiled-kernel cache of the library (cubins keyed by generated source + kernel headers + NVRTC version):
# kept inside the package's own build/ tree unless the caller chose a place
_CACHE_DIR = os.environ.setdefault("SDE_B200_CACHE", os.path.join(os.path.dirname(_HERE), "build", "jit_cache"))
try:
    os.makedirs(_CACHE_DIR, exist_ok=True)
except OSError:
    pass


def lib():
    """Load libsde_b200.so (fails loudly when the native library has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python sde-sim-rs_b200/build.py` (or __graft_entry__.build()). "
                "sde_sim_rs has no pure-Python or CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().sde_last_error() or b"").decode(errors="replace")


def check(rc: int, prefix_value: str = "", prefix_runtime: str = "") -> None:
    """Map return codes the way the pyo3 layer maps Rust errors (src/py_binding.rs:20-53)."""
    if rc == SDE_OK:
        return
    msg = last_error()
    if rc == SDE_ERR_VALUE:
        raise ValueError(prefix_value + msg)
    raise RuntimeError(prefix_runtime + msg)


def cstr_array(strs):
    return (C.c_char_p * len(strs))(*[s.encode() for s in strs])


def default_options() -> SdeOptions:
    o = SdeOptions()
    lib().sde_options_default(C.byref(o))
    return o
