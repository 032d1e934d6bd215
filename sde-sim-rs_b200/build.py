#!/usr/bin/env python3
"""Builds libsde_b200.so in-tree (sde-sim-rs_b200/sde_sim_rs/libsde_b200.so).

Steps (all cross-compile without a GPU):
  1. nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo  csrc/kernels/sde_util_kernels.cu
  2. nvcc syntax/ptxas check of the fused kernel template against a representative generated
     model (the same translation unit NVRTC compiles at plan creation)
  3. g++: host sources + blobs.S (kernel headers for NVRTC, util cubin, Joe-Kuo table) -> shared library
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
KDIR = os.path.join(CSRC, "kernels")
HDIR = os.path.join(CSRC, "host")
BUILD = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "sde_sim_rs", "libsde_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp/libstdc++ paths for shared links
    for c in ("/usr/bin/g++", shutil.which("g++") or "g++"):
        if os.path.exists(c):
            return c
    return "g++"


def run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def write_blobs(kernel_headers: list[str], util_cubin: str, joe_kuo: str) -> str:
    """Assembler file that .incbin's the payloads linked into the library: the kernel headers
    NVRTC compiles against, the ahead-of-time util cubin and the Joe-Kuo parameter table."""
    items = [(os.path.basename(h).replace("sde_", "").replace(".", "_"), h) for h in kernel_headers]
    items += [("util_cubin", util_cubin), ("joe_kuo", joe_kuo)]
    lines = ["    .section .rodata"]
    for sym, path in items:
        lines += [f"    .global sde_blob_{sym}_begin", f"    .global sde_blob_{sym}_end", "    .balign 16",
                  f"sde_blob_{sym}_begin:", f'    .incbin "{path}"', f"sde_blob_{sym}_end:", "    .byte 0"]
    lines.append('    .section .note.GNU-stack,"",@progbits')
    out = os.path.join(BUILD, "blobs.S")
    text = "\n".join(lines) + "\n"
    if not os.path.exists(out) or open(out).read() != text:
        with open(out, "w") as f:
            f.write(text)
    return out


def build(force: bool = False, verbose_ptxas: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    kernel_headers = [os.path.join(KDIR, f) for f in sorted(os.listdir(KDIR)) if f.endswith(".cuh")]
    util_cu = os.path.join(KDIR, "sde_util_kernels.cu")
    util_cubin = os.path.join(BUILD, "sde_util_kernels.cubin")
    if force or newer(util_cubin, kernel_headers + [util_cu]):
        run([NVCC, "-cubin", *GENCODE, "-lineinfo", "-O3", "-std=c++17", "-I", KDIR,
             *(["-Xptxas", "-v"] if verbose_ptxas else []), "-o", util_cubin, util_cu])
    host_srcs = [os.path.join(HDIR, f) for f in sorted(os.listdir(HDIR)) if f.endswith(".cpp")]
    host_hdrs = [os.path.join(HDIR, f) for f in sorted(os.listdir(HDIR)) if f.endswith(".h")]
    jk = os.path.join(HERE, "data", "joe_kuo_d6_21201.bin")
    blobs = write_blobs(kernel_headers, util_cubin, jk)
    inc = os.path.join(HERE, "..", "include", "sde_b200.h")
    if force or newer(OUT, host_srcs + host_hdrs + kernel_headers + [blobs, jk, util_cubin, inc, __file__]):
        cxx = _cxx()
        run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-Wno-unused-parameter",
             "-fvisibility=hidden", "-I", os.path.join(CUDA_HOME, "include"), "-I", HDIR,
             *host_srcs, blobs, "-ldl", "-lpthread", "-o", OUT])
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
    print(OUT)
