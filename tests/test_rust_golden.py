"""Parity against golden vectors produced by the UNMODIFIED reference crate (tools/rust_golden).

The fixture `tests/golden/rust_v1.json` is written by `cargo run --release` in tools/rust_golden on a machine with a Rust
toolchain (this repository's build image has none: no cargo / rustc, no network for the crates).  While the file is
absent every test here SKIPS with that reason and parity stays "unpinned" (DESIGN.md §7); the day it lands, the CPU
tests pin the oracle to the reference and the `-m gpu` tests pin the CUDA path to it, with no code change.

Bars: bit-exact for the ChaCha8 f64 stream, raw Sobol points, shifted uniforms, Poisson counts and the parser table;
<= 2 ulp for the A&S inverse normal and expression values (libm `ln` / `sin` / `pow` versions may differ by an ulp);
<= 1e-12 relative for whole paths (north_star); the number of bit-identical values is printed.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

import sde_sim_rs as S

FIXTURE = os.path.join(ROOT, "tests", "golden", "rust_v1.json")
_WHY = ("tests/golden/rust_v1.json is absent: it is produced by tools/rust_golden (cargo run --release) from the unmodified "
        "reference crate; no Rust toolchain exists in this image — PARITY STAYS UNPINNED until the fixture is committed")


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(FIXTURE):
        pytest.skip(_WHY)
    with open(FIXTURE) as f:
        g = json.load(f)
    assert g.get("format") == "sde-golden-1"
    return g


def f64(bits):
    return np.asarray(bits, dtype=np.uint64).view(np.float64)


def ulp_diff(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.spacing(np.abs(b)), 5e-324)
    d[both_nan] = 0.0
    d[np.isnan(d)] = np.inf
    return d


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _case_inputs(c):
    times = f64(c["times_bits"])
    init = {k: float(f64([v])[0]) for k, v in c["init_bits"].items()}
    return c["equations"], times, init


# ---------------------------------------------------------------- CPU: the oracle against the reference
def test_oracle_chacha8_f64_stream(gold, oracle):
    for c in gold["pseudo_f64"]:
        want = f64(c["bits"])
        assert np.array_equal(oracle.chacha8_f64(int(c["seed"]), want.size), want), c["seed"]      # src/rng/pseudo.rs:14-31


def test_oracle_sobol_raw_points(gold, oracle):
    for c in gold["sobol_raw"]:
        want = np.stack([f64(r) for r in c["bits"]])
        V = oracle.sobol_direction_numbers(int(c["dims"]))
        got = oracle.sobol_points(V, int(c["first_point_index"]), want.shape[0]).astype(np.float64) * 2.0**-64
        assert np.array_equal(got, want), c["dims"]                                               # src/rng/sobol.rs:15-25


def test_oracle_sobol_shifted_uniforms(gold, oracle):
    for c in gold["sobol_shifted"]:
        K, T, n = int(c["K"]), int(c["T"]), int(c["n_paths"])
        eqs = ["dA = ( 1.0 ) * dW1 + ( 1.0 ) * dW2"] if K == 2 else ["dA = ( 1.0 ) * dW1"]
        U = oracle.Universe(eqs, np.arange(T, dtype=np.float64))
        got = oracle.uniforms(U, n, "sobol", seed=int(c["seed"]), scramble="cp_shift_per_path")
        want = np.stack([f64(r) for r in c["bits"]]).reshape(n, T - 1, K)
        assert np.array_equal(got, want)                                                          # src/rng/sobol.rs:35-53,62-79


def test_oracle_inverse_cdfs(gold, oracle):
    p, z = f64(gold["icdf_normal"]["p_bits"]), f64(gold["icdf_normal"]["z_bits"])
    d = ulp_diff(oracle.icdf_normal(p), z)
    print("icdf normal: bit-identical", int((d == 0).sum()), "of", d.size, "max ulp", d.max())
    assert d.max() <= 2
    u, lam, k = (f64(gold["icdf_poisson"][key]) for key in ("u_bits", "lambda_bits", "k_bits"))
    got = np.array([oracle.icdf_poisson(float(a), float(b)) for a, b in zip(u, lam)], dtype=np.float64)
    assert np.array_equal(got, k)


def test_oracle_expressions(gold, oracle):
    e = gold["expr"]
    t = float(f64([e["t_bits"]])[0])
    vars_ = {k: float(f64([v])[0]) for k, v in e["vars"].items()}
    worst = 0.0
    for src, want in zip(e["src"], f64(e["value_bits"])):
        got = oracle.expr_eval(src, vars_, t)
        d = float(ulp_diff([got], [want])[0])
        worst = max(worst, d)
        assert d <= 2, (src, got, want)                                                           # fasteval 0.2.4 semantics
    print("expressions: worst ulp distance", worst)


def test_parser_table(gold, oracle):
    for c in gold["parser"]:
        for make in (lambda eq: oracle.Universe(eq, [0.0, 0.5, 1.0]), lambda eq: S.Universe(eq, [0.0, 0.5, 1.0])):
            if not c["ok"]:
                with pytest.raises(ValueError):
                    make(c["equations"])
                continue
            u = make(c["equations"])
            names = getattr(u, "names", None) or u.process_names
            factors = getattr(u, "factors", None)
            factors = u.factor_names if factors is None else factors
            assert (list(names), [bool(b) for b in u.is_levy], [int(n) for n in u.num_terms], list(factors)) == \
                   (c["names"], c["is_levy"], c["num_terms"], c["factors"]), c["equations"]


def test_oracle_paths(gold, oracle):
    for c in gold["paths"]:
        eqs, times, init = _case_inputs(c)
        U = oracle.Universe(eqs, times)
        got = oracle.simulate(U, init, int(c["n_paths"]), c["scheme"], c["rng_method"], seed=int(c["seed"]),
                              scramble="cp_shift_per_path", nthreads=1)
        want = f64(c["values_bits"]).reshape(got.shape)
        same = int((got == want).sum())
        print(c["name"], "bit-identical", same, "of", want.size, "max rel", rel_err(got, want))
        assert rel_err(got, want) <= 1e-12, c["name"]


def test_oracle_trace(gold, oracle):
    eqs, times = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"], [k / 252 for k in range(4)]
    for c in gold["trace"]:
        u = f64(c["u_bits"])
        U = oracle.Universe(eqs, times)
        inj = np.zeros((1, 3, 2))
        inj[0, :, 0] = oracle.icdf_normal(u)
        inj[0, :, 1] = u
        got = oracle.simulate(U, {"X1": 1.0}, 1, c["scheme"], inject=inj)
        assert rel_err(got.ravel(), f64(c["values_bits"])) <= 1e-13, c["scheme"]


# ---------------------------------------------------------------- GPU: the CUDA path against the reference
@pytest.mark.gpu
def test_device_chacha_and_sobol(gold):
    import ctypes as C

    from sde_sim_rs import _ffi

    for c in gold["pseudo_f64"]:
        want = f64(c["bits"])
        out = np.zeros(want.size, dtype=np.uint64)
        _ffi.check(_ffi.lib().sde_chacha8_u64(0, int(c["seed"]), out.size, out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal((out >> np.uint64(11)).astype(np.float64) * 2.0**-53, want)
    for c in gold["sobol_raw"]:
        want = np.stack([f64(r) for r in c["bits"]])
        out = np.zeros(want.shape, dtype=np.uint64)
        _ffi.check(_ffi.lib().sde_sobol_points(0, int(c["dims"]), int(c["first_point_index"]), want.shape[0], out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(out.astype(np.float64) * 2.0**-64, want)


@pytest.mark.gpu
def test_device_paths(gold):
    for c in gold["paths"]:
        eqs, times, init = _case_inputs(c)
        got = S.simulate(eqs, times, int(c["n_paths"]), init, c["rng_method"], c["scheme"], seed=int(c["seed"])).to_numpy()
        want = f64(c["values_bits"]).reshape(got.shape)
        flips = 0
        if any("dN" in e for e in eqs):
            # a Poisson count can flip where u sits within an ulp of a CDF step (CUDA exp vs libm exp): compare per path
            ok = np.array([rel_err(got[i], want[i]) <= 1e-12 for i in range(got.shape[0])])
            flips = int((~ok).sum())
            assert flips <= 1, (c["name"], flips)
            got, want = got[ok], want[ok]
        print(c["name"], "max rel", rel_err(got, want), "paths with a flipped jump", flips)
        assert rel_err(got, want) <= 1e-12, c["name"]
