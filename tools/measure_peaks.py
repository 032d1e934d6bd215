#!/usr/bin/env python3
"""Device microbenchmarks recorded beside MEASURED_PEAKS.json: pure-write fill GB/s,
DFMA and FFMA issue peaks, measured by libsde_b200's own kernels (csrc/kernels/sde_util_kernels.cu)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sde-sim-rs_b200"))
from sde_sim_rs import _ffi  # noqa: E402

fill, dfma, ffma = C.c_double(), C.c_double(), C.c_double()
_ffi.check(_ffi.lib().sde_measure_peaks(0, C.byref(fill), C.byref(dfma), C.byref(ffma)))
res = {"fill_gbs": fill.value, "dfma_tflops": dfma.value, "ffma_tflops": ffma.value,
       "how": "sde_k_fill: 8 GiB of 16-byte stores, best of 10; sde_k_dfma / sde_k_ffma: 8 independent FMA chains per thread, 148*8 CTAs x 256 threads, best of 5 (2 flop per FMA)"}
print(json.dumps(res))
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
with open(os.path.join(out, "device_peaks.json"), "w") as f:
    json.dump(res, f, indent=1)
