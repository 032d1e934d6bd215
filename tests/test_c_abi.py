"""The C-ABI from plain C: tests/c_abi/consumer.c is compiled with gcc -std=c99 against include/sde_b200.h and linked to
the in-tree libsde_b200.so — what a Rust `-sys` crate / cgo / JNI stub would do.  Without a GPU the
simulate call must fail loudly (no CPU fallback); on a B200 its numbers must match the Python mirror and the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GBM_EQ, PKG, ROOT, grid

import sde_sim_rs as S

LIBDIR = os.path.join(PKG, "sde_sim_rs")
SRC = os.path.join(ROOT, "tests", "c_abi", "consumer.c")


@pytest.fixture(scope="module")
def consumer(tmp_path_factory):
    S._ffi.lib()                                             # makes sure the library is built
    exe = str(tmp_path_factory.mktemp("c_abi") / "consumer")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", LIBDIR, "-lsde_b200", f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run(exe, *args):
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=300)


def test_header_is_c99_and_library_links_from_c(consumer):
    r = _run(consumer, 64, 8, 1, "pseudo", "euler")
    assert "bad equation: rc=1" in r.stdout                  # SDE_ERR_VALUE, like the pyo3 layer's ValueError
    assert "universe: P=1 K=1 T=9 name=X1 factor=dW1" in r.stdout
    if not S.cuda_available():
        assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 0 and "checks: ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("rng,scheme", [("pseudo", "euler"), ("sobol", "euler"), ("pseudo", "runge-kutta")])
def test_c_caller_matches_python_mirror_and_oracle(consumer, oracle, rng, scheme):
    N, D, seed = 500, 32, 7
    r = _run(consumer, N, D, seed, rng, scheme)
    assert r.returncode == 0 and "checks: ok" in r.stdout and "all checks: ok" in r.stdout, r.stdout + r.stderr
    assert "shard union bit-identical" in r.stdout and "merged moments ok" in r.stdout, r.stdout
    vals = {k: float(v) for k, v in re.findall(r"^(sum|terminal_mean|first_path_terminal)=(\S+)$", r.stdout, re.M)}
    kw = {"scramble": "xor"} if rng == "sobol" else {}
    got = S.simulate(GBM_EQ, grid(252, D), N, {"X1": 1.0}, rng, scheme, seed=seed, **kw).to_numpy()
    assert vals["first_path_terminal"] == got[0, -1, 0]      # same library, same call: bit-identical
    assert abs(vals["sum"] - got.sum()) <= 1e-12 * got.sum()
    ref = oracle.simulate(oracle.Universe(GBM_EQ, grid(252, D)), {"X1": 1.0}, N, scheme, rng, seed=seed, **kw)
    assert abs(vals["terminal_mean"] / ref[:, -1, 0].mean() - 1) <= 1e-12
    assert abs(vals["first_path_terminal"] / ref[0, -1, 0] - 1) <= 1e-12
