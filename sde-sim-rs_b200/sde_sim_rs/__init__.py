"""sde_sim_rs — B200-native drop-in for the simulation hot path of Aschii85/sde-sim-rs.

Mirrors the reference's Python surface (python/sde_sim_rs/__init__.py:1-3 re-exporting the
pyo3 function src/py_binding.rs:8-56):

    simulate(processes_equations, time_steps, scenarios, initial_values, rng_method, scheme)

Same argument names, meaning and errors (ValueError for scenarios <= 0 and unparsable
equations, RuntimeError for simulation failures).  The reference returns a polars
DataFrame in long format; polars is not importable here, so `simulate` returns a
`Filtration` that holds the dense value tensor on the GPU and renders the same long-format
columns (scenario:i32, time:f64, process_name:str, value:f64; src/filtration.rs:108-113) on
request (`.columns()`, `.to_pandas()`, `.to_polars()` when polars exists).

Everything runs through libsde_b200.so (hand-written sm_100a CUDA behind a C-ABI).  PyTorch is
used only for device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import (ARITH_FAST, ARITH_STRICT, DTYPE_F32, DTYPE_F64, ICDF_FAST, ICDF_REFERENCE, ICDF_SINGLE, LAYOUT_NTP, LAYOUT_TPN, OUT_MOMENTS,
                   OUT_PATHS, OUT_TERMINAL, RK_REFERENCE, RK_TEXTBOOK, SCRAMBLE_CP_SHIFT_PER_PATH, SCRAMBLE_NONE,
                   SCRAMBLE_XOR)

__all__ = ["simulate", "simulate_frame", "simulate_devices", "simulate_sharded", "DevicePlans", "merge_moments_device", "parse_equations", "Universe", "Plan", "Filtration", "shard_range", "merge_moments",
           "cuda_available", "version"]

_OUTPUTS = {"paths": OUT_PATHS, "terminal": OUT_TERMINAL, "moments": OUT_MOMENTS}
_LAYOUTS = {"NTP": LAYOUT_NTP, "TPN": LAYOUT_TPN}
_SCRAMBLES = {"cp_shift_per_path": SCRAMBLE_CP_SHIFT_PER_PATH, "xor": SCRAMBLE_XOR, "none": SCRAMBLE_NONE}
_DTYPES = {"f64": DTYPE_F64, "float64": DTYPE_F64, "f32": DTYPE_F32, "float32": DTYPE_F32}
_ICDFS = {"reference": ICDF_REFERENCE, "fast": ICDF_FAST, "single": ICDF_SINGLE}
_ARITHS = {"strict": ARITH_STRICT, "fast": ARITH_FAST}
_RKS = {"reference": RK_REFERENCE, "textbook": RK_TEXTBOOK}
_GENERATORS = {"chacha8": _ffi.GEN_CHACHA8, "philox": _ffi.GEN_PHILOX}


def version() -> str:
    return _ffi.lib().sde_version().decode()


def cuda_available() -> bool:
    return bool(_ffi.lib().sde_cuda_available())


def _pick(table: dict, key: str, what: str) -> int:
    try:
        return table[key]
    except KeyError:
        raise ValueError(f"unknown {what} {key!r}; expected one of {sorted(table)}") from None


class Universe:
    """proc::ProcessUniverse (src/proc/mod.rs:61-90) built by proc::util::parse_equations."""

    def __init__(self, processes_equations: Sequence[str], time_steps: Sequence[float]):
        L = _ffi.lib()
        self.equations = [str(e) for e in processes_equations]
        self.time_steps = np.ascontiguousarray(np.asarray(time_steps, dtype=np.float64))
        h = C.c_void_p()
        rc = L.sde_universe_parse(_ffi.cstr_array(self.equations), len(self.equations),
                                  self.time_steps.ctypes.data_as(C.c_void_p), self.time_steps.size, C.byref(h))
        _ffi.check(rc, prefix_value="Failed to parse equations: ")          # py_binding.rs:30-32
        self._h = h
        self.num_processes = L.sde_universe_num_processes(h)
        self.num_factors = L.sde_universe_num_factors(h)
        self.process_names = [L.sde_universe_process_name(h, i).decode() for i in range(self.num_processes)]
        self.is_levy = [bool(L.sde_universe_process_is_levy(h, i)) for i in range(self.num_processes)]
        self.num_terms = [L.sde_universe_process_num_terms(h, i) for i in range(self.num_processes)]
        self.factor_names = [L.sde_universe_factor_name(h, k).decode() for k in range(self.num_factors)]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _ffi is not None and getattr(_ffi, "_lib", None) is not None:     # may run at interpreter shutdown
            _ffi._lib.sde_universe_free(h)


def parse_equations(processes_equations: Sequence[str], time_steps: Sequence[float]) -> Universe:
    """proc::util::parse_equations (src/proc/util.rs:52-66)."""
    return Universe(processes_equations, time_steps)


def _make_options(*, device: int, seed: int, scenario_offset: int, output: str, layout: str, scramble: str, icdf: str,
                  arithmetic: str, rk_variant: str, inject_ptr: int = 0, tile_steps: int = 0, block_threads: int = 0,
                  min_blocks: int = 0, ntp_direct: int = 0, dtype: str = "f64", wide_mma: int = 0, generator: str = "chacha8"):
    o = _ffi.default_options()
    o.device = device
    o.seed = seed & (2**64 - 1)
    o.scenario_offset = scenario_offset
    o.output = _pick(_OUTPUTS, output, "output")
    o.layout = _pick(_LAYOUTS, layout, "layout")
    o.scramble = _pick(_SCRAMBLES, scramble, "scramble")
    o.icdf = _pick(_ICDFS, icdf, "icdf")
    o.arith = _pick(_ARITHS, arithmetic, "arithmetic")
    o.rk_variant = _pick(_RKS, rk_variant, "rk_variant")
    o.inject = inject_ptr or None
    o.tile_steps = tile_steps
    o.block_threads = block_threads
    o.min_blocks = min_blocks
    o.ntp_direct = ntp_direct
    o.dtype = _pick(_DTYPES, dtype, "dtype")
    o.wide_mma = wide_mma
    o.generator = _pick(_GENERATORS, generator, "generator")
    return o


def _current_device() -> int:
    import torch

    return torch.cuda.current_device() if torch.cuda.is_available() else 0


class Plan:
    """One model lowered to sm_100a device code for one (scheme, rng_method, options) choice."""

    def __init__(self, universe: Universe, scheme: str = "euler", rng_method: str = "pseudo", *, output: str = "paths",
                 layout: str = "NTP", scramble: str = "cp_shift_per_path", icdf: str = "reference",
                 arithmetic: str = "strict", rk_variant: str = "reference", device: Optional[int] = None,
                 inject=None, tile_steps: int = 0, block_threads: int = 0, min_blocks: int = 0, ntp_direct: int = 0,
                 dtype: str = "f64", wide_mma: int = 0, generator: str = "chacha8"):
        self.universe = universe
        self.dtype = "f32" if _pick(_DTYPES, dtype, "dtype") == DTYPE_F32 else "f64"
        self.scheme, self.rng_method = scheme, rng_method
        self.output, self.layout = output, layout
        self.device = _current_device() if device is None else int(device)
        self._inject = inject                                   # keep the tensor alive
        opts = _make_options(device=self.device, seed=0, scenario_offset=0, output=output, layout=layout,
                             scramble=scramble, icdf=icdf, arithmetic=arithmetic, rk_variant=rk_variant,
                             inject_ptr=(inject.data_ptr() if inject is not None else 0), tile_steps=tile_steps,
                             block_threads=block_threads, min_blocks=min_blocks, ntp_direct=ntp_direct, dtype=dtype,
                             wide_mma=wide_mma, generator=generator)
        h = C.c_void_p()
        rc = _ffi.lib().sde_plan_create(universe._h, scheme.encode(), rng_method.encode(), C.byref(opts), C.byref(h))
        _ffi.check(rc, prefix_runtime="Simulation failed: ")
        self._h = h
        self.launches = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _ffi is not None and getattr(_ffi, "_lib", None) is not None:
            _ffi._lib.sde_plan_free(h)

    @property
    def source(self) -> str:
        return _ffi.lib().sde_plan_source(self._h).decode()

    @property
    def prelowered(self) -> bool:
        """True when the cubin came from the ahead-of-time cache on disk instead of an NVRTC compile at plan creation."""
        return bool(_ffi.lib().sde_plan_is_prelowered(self._h))

    def output_shape(self, scenarios: int):
        T, P = self.universe.time_steps.size, self.universe.num_processes
        if self.output == "paths":
            return (scenarios, T, P) if self.layout == "NTP" else (T, P, scenarios)
        if self.output == "terminal":
            return (scenarios, P)
        return (P, 3)

    @staticmethod
    def _init_arrays(initial_values: Dict[str, float]):
        names = list(initial_values)
        vals = np.asarray([float(initial_values[k]) for k in names], dtype=np.float64)
        return _ffi.cstr_array(names), vals, len(names)

    def run(self, initial_values: Dict[str, float], scenarios: int, *, seed: int = 0, scenario_offset: int = 0,
            out=None, stream=None):
        """Launch on the current torch stream; returns a CUDA float64 tensor (device resident)."""
        import torch

        if scenarios <= 0:
            raise ValueError("scenarios must be a positive integer")
        shape = self.output_shape(scenarios)
        tdt = torch.float32 if (self.dtype == "f32" and self.output != "moments") else torch.float64
        if out is None:
            out = torch.empty(shape, dtype=tdt, device=f"cuda:{self.device}")
        elif tuple(out.shape) != tuple(shape) or out.dtype != tdt or not out.is_cuda or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous CUDA {tdt} tensor of shape {shape}")
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        names, vals, n = self._init_arrays(initial_values)
        nl = C.c_int(0)
        rc = _ffi.lib().sde_plan_run_device(self._h, names, vals.ctypes.data_as(C.c_void_p), n, scenarios,
                                            seed & (2**64 - 1), scenario_offset, C.c_void_p(out.data_ptr()),
                                            C.c_void_p(stream), C.byref(nl))
        _ffi.check(rc, prefix_runtime="Simulation failed: ")
        self.launches += nl.value
        return out

    def run_host(self, initial_values: Dict[str, float], scenarios: int, *, seed: int = 0, scenario_offset: int = 0,
                 out=None):
        """HOST buffers end to end: device memory managed by the library, chunked D2H overlapped with compute.

        `out` may be a numpy array or a (pinned) CPU torch tensor; returns it."""
        if scenarios <= 0:
            raise ValueError("scenarios must be a positive integer")
        shape = self.output_shape(scenarios)
        ndt = np.float32 if (self.dtype == "f32" and self.output != "moments") else np.float64
        if out is None:
            out = np.empty(shape, dtype=ndt)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        isz = out.element_size() if hasattr(out, "element_size") else out.itemsize
        if int(np.prod(out.shape)) != int(np.prod(shape)) or isz != np.dtype(ndt).itemsize:
            raise ValueError(f"out must hold {shape} {np.dtype(ndt).name} values")
        names, vals, n = self._init_arrays(initial_values)
        nl = C.c_int(0)
        rc = _ffi.lib().sde_plan_run_host(self._h, names, vals.ctypes.data_as(C.c_void_p), n, scenarios,
                                          seed & (2**64 - 1), scenario_offset, C.c_void_p(ptr), C.byref(nl))
        _ffi.check(rc, prefix_runtime="Simulation failed: ")
        self.launches += nl.value
        return out


class Filtration:
    """Result of `simulate`: the value column of every ScenarioFiltration (src/filtration.rs:12-19,112),
    dense on the GPU, with the reference's long-format view built on request."""

    def __init__(self, values, time_steps: np.ndarray, process_names: List[str], *, output: str, layout: str,
                 scenario_offset: int = 0, seed: int = 0):
        self.values = values                # torch.cuda tensor
        self.time_steps = time_steps
        self.process_names = process_names
        self.output, self.layout = output, layout
        self.scenario_offset = scenario_offset
        self.seed = seed

    @property
    def shape(self):
        return tuple(self.values.shape)

    def dense(self):
        """[N, T, P] view (reference row order) for full-path results."""
        if self.output != "paths":
            raise ValueError("dense() needs output='paths'")
        return self.values if self.layout == "NTP" else self.values.permute(2, 0, 1)

    def to_numpy(self) -> np.ndarray:
        return self.values.detach().cpu().numpy()

    def moments(self) -> Dict[str, Dict[str, float]]:
        if self.output != "moments":
            raise ValueError("moments() needs output='moments'")
        m = self.to_numpy()
        return {name: {"count": float(m[i, 0]), "mean": float(m[i, 1]), "m2": float(m[i, 2]),
                       "variance": float(m[i, 2] / (m[i, 0] - 1)) if m[i, 0] > 1 else float("nan")}
                for i, name in enumerate(self.process_names)}

    def columns(self) -> Dict[str, np.ndarray]:
        """The four columns of ScenarioFiltration::to_lazyframe concatenated over scenarios
        (src/filtration.rs:87-113, src/sim/mod.rs:88-91): rows ordered (scenario, time, process)."""
        dense = self.dense().detach().cpu().numpy()
        N, T, P = dense.shape
        scen = (np.arange(N, dtype=np.int64) + self.scenario_offset).astype(np.int32)     # `s_idx as i32` wraps (sim/mod.rs:47)
        return {
            "scenario": np.repeat(scen, T * P),
            "time": np.tile(np.repeat(self.time_steps, P), N),
            "process_name": np.tile(np.asarray(self.process_names, dtype=object), N * T),
            "value": dense.reshape(-1),
        }

    def to_pandas(self):
        """Long frame with the reference's columns and dtypes.  Built through Arrow when pyarrow is importable: the string
        column is then dictionary-expanded in native code (20 M rows: 1.4 s instead of 5 s through a Python object array)."""
        import pandas as pd

        try:
            return self.to_arrow().to_pandas()
        except ImportError:
            return pd.DataFrame(self.columns())

    def to_arrow(self):
        """Arrow table with the reference's schema (scenario:int32, time:float64, process_name:string, value:float64);
        the value column wraps the host copy of the dense buffer without another copy.  pyo3-polars hands the
        reference's frame to Python as exactly these Arrow buffers (src/py_binding.rs:55)."""
        import pyarrow as pa

        c = self.columns()
        names = pa.DictionaryArray.from_arrays(
            pa.array(np.tile(np.arange(len(self.process_names), dtype=np.int32), c["value"].size // len(self.process_names))),
            pa.array(self.process_names)).cast(pa.string())
        return pa.table({"scenario": pa.array(c["scenario"], type=pa.int32()), "time": pa.array(c["time"]),
                         "process_name": names, "value": pa.array(c["value"])})

    def to_polars(self):
        import polars as pl  # not in this image; present wherever the reference's callers run

        c = self.columns()
        return pl.DataFrame({"scenario": c["scenario"], "time": c["time"],
                             "process_name": [str(x) for x in c["process_name"]], "value": c["value"]})


_PLAN_CACHE: Dict[tuple, Plan] = {}


def _cached_plan(equations, time_steps, scheme, rng_method, **kw) -> Plan:
    ts = np.ascontiguousarray(np.asarray(time_steps, dtype=np.float64))
    key = (tuple(equations), ts.tobytes(), scheme, rng_method, tuple(sorted(kw.items())))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        plan = Plan(Universe(equations, ts), scheme, rng_method, **kw)
        if len(_PLAN_CACHE) > 64:
            _PLAN_CACHE.clear()
        _PLAN_CACHE[key] = plan
    return plan


_EXTENSIONS = ("seed", "output", "layout", "scramble", "icdf", "arithmetic", "rk_variant", "compat", "device", "devices", "scenario_offset",
               "dtype", "generator")


def simulate(processes_equations: Sequence[str], time_steps: Sequence[float], scenarios: int,
             initial_values: Dict[str, float], rng_method: str = "pseudo", scheme: str = "euler", *,
             frame="auto", **ext):
    """Drop-in for sde_sim_rs.simulate (src/py_binding.rs:10-18; defaults as in python/sde_sim_rs/sde_sim_rs.pyi:11-12).

    Called like the reference — the six reference arguments only — it returns what the reference returns: the long
    DataFrame `scenario:i32, time:f64, process_name:str, value:f64` in (scenario, time, process) order
    (src/py_binding.rs:51-55, src/filtration.rs:108-113): a polars frame where polars is importable, else pandas.

    Keyword-only extensions: `seed` (the reference draws a fresh OS-entropy seed per call, src/sim/mod.rs:28-29 — so does
    this when seed is None), `output` paths|terminal|moments, `layout`, `scramble` cp_shift_per_path (reference
    behaviour) | xor | none, `icdf` reference|fast|single, `arithmetic` strict|fast, `rk_variant` reference|textbook,
    `device`, `scenario_offset`, `dtype` f64|f32 (f32: state, arithmetic and stored values in single precision; needs
    arithmetic="fast"), `generator` chacha8 (the reference's stream, bit-exact) | philox (Philox4x32-10: a cheaper
    counter-based stream for rng_method != "sobol"; agrees with the reference statistically, not draw for draw),
    `compat` (SURVEY's name for `rk_variant`: "reference" | "textbook"), `devices` (a list of GPU ordinals, or "all": the
    scenarios are sharded over them inside one C call like rayon's par_iter, see `simulate_devices` — the result is then the
    list of per-device `Filtration` shards, or one `Filtration` of merged moments).  With any of them the result is a
    `Filtration` (the dense value tensor, resident on the GPU) unless frame=True; frame=False always returns the `Filtration`.
    """
    unknown = [k for k in ext if k not in _EXTENSIONS]
    if unknown:
        raise TypeError(f"simulate() got unexpected keyword argument(s) {unknown}; extensions are {list(_EXTENSIONS)}")
    if frame not in (True, False, "auto"):
        raise ValueError("frame must be True, False or 'auto'")
    want_frame = (len(ext) == 0) if frame == "auto" else bool(frame)
    if "compat" in ext:
        compat = ext.pop("compat")
        if "rk_variant" in ext and ext["rk_variant"] != compat:
            raise ValueError("compat and rk_variant name the same option and disagree")
        ext["rk_variant"] = compat
    if ext.get("devices") is not None:
        devices = ext.pop("devices")
        if ext.pop("device", None) is not None:
            raise ValueError("give either device or devices")
        if ext.pop("scenario_offset", 0):
            raise ValueError("scenario_offset is per shard when devices is given")
        shards = simulate_devices(processes_equations, time_steps, scenarios, initial_values, rng_method, scheme,
                                  devices=None if devices == "all" else devices, **ext)
        if not want_frame:
            return shards
        if isinstance(shards, Filtration):
            raise ValueError("frame=True needs output='paths' (the reference's frame holds every row)")
        try:
            import polars as pl
            return pl.concat([f.to_polars() for f in shards])
        except ImportError:
            import pandas as pd
            return pd.concat([f.to_pandas() for f in shards], ignore_index=True)
    ext.pop("devices", None)
    res = _simulate_filtration(processes_equations, time_steps, scenarios, initial_values, rng_method, scheme, **ext)
    if not want_frame:
        return res
    if res.output != "paths":
        raise ValueError("frame=True needs output='paths' (the reference's frame holds every row)")
    try:
        return res.to_polars()
    except ImportError:
        return res.to_pandas()


def _simulate_filtration(processes_equations: Sequence[str], time_steps: Sequence[float], scenarios: int,
                         initial_values: Dict[str, float], rng_method: str = "pseudo", scheme: str = "euler", *,
                         seed: Optional[int] = None, output: str = "paths", layout: str = "NTP",
                         scramble: str = "cp_shift_per_path", icdf: str = "reference", arithmetic: str = "strict",
                         rk_variant: str = "reference", device: Optional[int] = None, scenario_offset: int = 0,
                         dtype: str = "f64", generator: str = "chacha8") -> Filtration:
    if not isinstance(scenarios, (int, np.integer)) or scenarios <= 0:
        raise ValueError("scenarios must be a positive integer")                      # py_binding.rs:20-24
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    # parse first so that equation errors surface as ValueError before any CUDA work (py_binding.rs:30-32)
    plan = _cached_plan(list(processes_equations), time_steps, scheme, rng_method, output=output, layout=layout,
                        scramble=scramble, icdf=icdf, arithmetic=arithmetic, rk_variant=rk_variant, device=device,
                        dtype=dtype, generator=generator)
    values = plan.run(dict(initial_values), int(scenarios), seed=seed, scenario_offset=scenario_offset)
    return Filtration(values, plan.universe.time_steps, plan.universe.process_names, output=output, layout=layout,
                      scenario_offset=scenario_offset, seed=seed)


def simulate_frame(processes_equations: Sequence[str], time_steps: Sequence[float], scenarios: int,
                   initial_values: Dict[str, float], rng_method: str = "pseudo", scheme: str = "euler", **kw):
    """The reference's call with the reference's return value: the long DataFrame `scenario:i32, time:f64,
    process_name:str, value:f64` in (scenario, time, process) row order (src/py_binding.rs:10-55,
    src/filtration.rs:108-113) — a polars frame where polars is importable (what pyo3-polars hands back), a pandas
    frame otherwise.  `simulate` itself returns the dense GPU tensor wrapped in a `Filtration`; this is the
    convenience for scripts written against the reference package."""
    return simulate(processes_equations, time_steps, scenarios, initial_values, rng_method, scheme, frame=True, **kw)


# ---------------------------------------------------------------- multi-GPU helpers
def shard_range(scenarios: int, rank: int, world_size: int):
    """Scenario range [lo, hi) of `rank`: contiguous, disjoint, union = [0, scenarios).  Sobol point
    indices and ChaCha keys are functions of the global scenario index, so the union of the shards is
    bit-identical to a single-GPU run."""
    base, rem = divmod(int(scenarios), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_moments(shards: np.ndarray) -> np.ndarray:
    """[n_shards, P, 3] (count, mean, M2) -> [P, 3], Chan et al. merge in shard order (deterministic)."""
    shards = np.ascontiguousarray(np.asarray(shards, dtype=np.float64))
    n, P, three = shards.shape
    assert three == 3
    out = np.zeros((P, 3), dtype=np.float64)
    _ffi.check(_ffi.lib().sde_moments_merge(shards.ctypes.data_as(C.c_void_p), n, P, out.ctypes.data_as(C.c_void_p)))
    return out


class DevicePlans:
    """One plan per GPU of `devices` behind ONE C call per run (sde_device_plans_create / sde_plan_run_devices): the
    host-side replacement of rayon's par_iter over scenarios (src/sim/mod.rs:41-43).  Moments are all-gathered over NCCL
    and Chan-merged on the devices inside that call; `collective` says how ("nccl", "peer" copies, or "none")."""

    def __init__(self, universe: Universe, scheme: str = "euler", rng_method: str = "pseudo", *, devices: Optional[Sequence[int]] = None,
                 output: str = "paths", layout: str = "NTP", scramble: str = "cp_shift_per_path", icdf: str = "reference",
                 arithmetic: str = "strict", rk_variant: str = "reference", dtype: str = "f64", block_threads: int = 0, wide_mma: int = 0,
                 generator: str = "chacha8"):
        import torch

        self.universe, self.output, self.layout = universe, output, layout
        self.dtype = "f32" if _pick(_DTYPES, dtype, "dtype") == DTYPE_F32 else "f64"
        devs = list(range(torch.cuda.device_count())) if devices is None else [int(d) for d in devices]
        if not devs:
            raise RuntimeError("Simulation failed: no CUDA device (there is no CPU fallback)")
        opts = _make_options(device=devs[0], seed=0, scenario_offset=0, output=output, layout=layout, scramble=scramble, icdf=icdf,
                             arithmetic=arithmetic, rk_variant=rk_variant, dtype=dtype, block_threads=block_threads, wide_mma=wide_mma,
                             generator=generator)
        h = C.c_void_p()
        arr = (C.c_int32 * len(devs))(*devs)
        rc = _ffi.lib().sde_device_plans_create(universe._h, scheme.encode(), rng_method.encode(), C.byref(opts), arr, len(devs), C.byref(h))
        _ffi.check(rc, prefix_runtime="Simulation failed: ")
        self._h, self.devices, self.launches, self.collective_ms = h, devs, 0, 0.0
        self.collective = {0: "none", 1: "nccl", 2: "peer"}[_ffi.lib().sde_device_plans_collective(h)]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _ffi is not None and getattr(_ffi, "_lib", None) is not None:
            _ffi._lib.sde_device_plans_free(h)

    def run(self, initial_values: Dict[str, float], scenarios: int, *, seed: int = 0, scenario_offset: int = 0):
        """Returns the per-device CUDA tensors: shard i of the paths / terminal values on device i, or (moments) the merged
        [P, 3] on every device.  Synchronous (every device has finished)."""
        import torch

        if scenarios <= 0:
            raise ValueError("scenarios must be a positive integer")
        T, P, G = self.universe.time_steps.size, self.universe.num_processes, len(self.devices)
        tdt = torch.float32 if (self.dtype == "f32" and self.output != "moments") else torch.float64
        outs = []
        for i, d in enumerate(self.devices):
            lo, hi = shard_range(scenarios, i, G)
            n = hi - lo
            shape = (P, 3) if self.output == "moments" else ((n, P) if self.output == "terminal" else ((n, T, P) if self.layout == "NTP" else (T, P, n)))
            outs.append(torch.empty(shape, dtype=tdt, device=f"cuda:{d}"))
        for d in self.devices:
            torch.cuda.synchronize(d)                            # the library launches on its own streams
        ptrs = (C.c_void_p * G)(*[t.data_ptr() if t.numel() else None for t in outs])
        names, vals, n_init = Plan._init_arrays(initial_values)
        nl, cms = C.c_int(0), C.c_double(0.0)
        rc = _ffi.lib().sde_plan_run_devices(self._h, names, vals.ctypes.data_as(C.c_void_p), n_init, scenarios, seed & (2**64 - 1),
                                             scenario_offset, ptrs, C.byref(nl), C.byref(cms))
        _ffi.check(rc, prefix_runtime="Simulation failed: ")
        self.launches += nl.value
        self.collective_ms = cms.value
        return outs


_DEVICE_PLAN_CACHE: Dict[tuple, DevicePlans] = {}


def simulate_devices(processes_equations, time_steps, scenarios, initial_values, rng_method="pseudo", scheme="euler", *,
                     devices: Optional[Sequence[int]] = None, seed: Optional[int] = None, output: str = "paths", **kw):
    """ONE process driving several GPUs of the box through one C call (sde_plan_run_devices; under torchrun use
    `simulate_sharded`): device i of `devices` (default: all visible) simulates `shard_range(scenarios, i, len(devices))` —
    disjoint Sobol index ranges / ChaCha keys, no data-path exchange; all launches are issued before any device is
    synchronised.  Returns the list of per-device `Filtration` shards (paths / terminal values stay on their GPU, in
    scenario order) or, for output="moments", one `Filtration` with the moments merged on the devices (NCCL all-gather
    + Chan-merge kernel)."""
    import torch

    if not isinstance(scenarios, (int, np.integer)) or scenarios <= 0:
        raise ValueError("scenarios must be a positive integer")
    devs = list(range(torch.cuda.device_count())) if devices is None else [int(d) for d in devices]
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    ts = np.ascontiguousarray(np.asarray(time_steps, dtype=np.float64))
    key = (tuple(processes_equations), ts.tobytes(), scheme, rng_method, tuple(devs), output, tuple(sorted(kw.items())))
    plans = _DEVICE_PLAN_CACHE.get(key)
    if plans is None:
        if len(_DEVICE_PLAN_CACHE) > 8:
            _DEVICE_PLAN_CACHE.clear()
        plans = _DEVICE_PLAN_CACHE[key] = DevicePlans(Universe(list(processes_equations), ts), scheme, rng_method, devices=devs, output=output, **kw)
    outs = plans.run(dict(initial_values), int(scenarios), seed=seed)
    uni, layout = plans.universe, kw.get("layout", "NTP")
    if output == "moments":
        return Filtration(outs[0], uni.time_steps, uni.process_names, output=output, layout="NTP", scenario_offset=0, seed=seed)
    shards = []
    for i, t in enumerate(outs):
        lo, hi = shard_range(int(scenarios), i, len(devs))
        if hi > lo:
            shards.append(Filtration(t, uni.time_steps, uni.process_names, output=output, layout=layout, scenario_offset=lo, seed=seed))
    return shards


def merge_moments_device(shards):
    """[n_shards, P, 3] CUDA tensor of (count, mean, M2) triples -> [P, 3] on the same device, by the library's merge
    kernel on the current stream (same arithmetic, bit for bit, as `merge_moments`)."""
    import torch

    assert shards.is_cuda and shards.dtype == torch.float64 and shards.dim() == 3 and shards.shape[2] == 3
    shards = shards.contiguous()
    out = torch.empty(shards.shape[1:], dtype=torch.float64, device=shards.device)
    dev = shards.device.index
    _ffi.check(_ffi.lib().sde_moments_merge_device(dev, C.c_void_p(shards.data_ptr()), shards.shape[0], shards.shape[1],
                                                   C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def simulate_sharded(processes_equations, time_steps, scenarios, initial_values, rng_method="pseudo", scheme="euler", *,
                     seed: int = 0, output: str = "moments", **kw) -> Filtration:
    """One process per GPU (torch.distributed already initialised): every rank simulates its scenario range;
    moments are all-gathered (3·P doubles per rank over NCCL/NVLink) and merged by the library's device kernel on the
    same stream — identically on every rank, no host hop; paths / terminal values stay resident on their GPU."""
    import torch
    import torch.distributed as dist

    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    if not isinstance(scenarios, (int, np.integer)) or scenarios <= 0:
        raise ValueError("scenarios must be a positive integer")                      # every rank raises: no collective entered
    lo, hi = shard_range(scenarios, rank, world)
    if hi > lo:
        res = _simulate_filtration(processes_equations, time_steps, hi - lo, initial_values, rng_method, scheme, seed=seed,
                                   output=output, scenario_offset=lo, **kw)
    else:
        # fewer scenarios than ranks: this rank owns none, but it still enters the collective with a zero-count triple
        uni = Universe(list(processes_equations), time_steps)
        T, P = uni.time_steps.size, uni.num_processes
        shape = {"moments": (P, 3), "terminal": (0, P)}.get(output, (0, T, P) if kw.get("layout", "NTP") == "NTP" else (T, P, 0))
        dev = kw.get("device")
        dev = _current_device() if dev is None else int(dev)
        dt = torch.float32 if (kw.get("dtype", "f64") in ("f32", "float32") and output != "moments") else torch.float64
        res = Filtration(torch.zeros(shape, dtype=dt, device=f"cuda:{dev}" if torch.cuda.is_available() else "cpu"),
                         uni.time_steps, uni.process_names, output=output, layout=kw.get("layout", "NTP"), scenario_offset=lo, seed=seed)
    if output == "moments" and world > 1:
        gathered = torch.empty((world,) + tuple(res.values.shape), dtype=res.values.dtype, device=res.values.device)
        dist.all_gather_into_tensor(gathered, res.values)
        res.values = merge_moments_device(gathered) if gathered.is_cuda else torch.from_numpy(merge_moments(gathered.numpy()))
    return res
