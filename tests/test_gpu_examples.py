"""The reference's own example models, verbatim (examples/example.py:10-19, examples/example.rs:11-28, examples/example_gbm.py:12-20),
as oracle-parity tests on the GPU, both schemes.  Scenario counts are reduced (the models are what is tested); the time grids are
the examples' own.  Bar: 1e-12 relative; paths that contain a jump may flip a Poisson count where u sits within an ulp of a CDF step
(CUDA exp vs glibc exp) and are compared per path."""
import numpy as np
import pytest

import sde_sim_rs as S

pytestmark = pytest.mark.gpu

EXAMPLE_PY = (["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dN1(X0)",
               "dX1 = ( 0.05 * X1 ) * dt + ( 0.2 * X1 ) * dW1 + ( 0.5 ) * dN1(X0)",
               "X2 = max(X1 - 100.0, 0.0)"], list(np.arange(0.0, 10.0, 0.01)), {"X0": 0.5, "X1": 100.0, "X2": 0.0}, "pseudo", "runge-kutta")
EXAMPLE_RS = (["dX1 = ( sin(t) ) * dt + (0.01 * X1) * dW1 + (0.001 * X1) * dN1(0.5 * cos(t))",
               "X2 = max(X1 - 100.0, 0.0)"], [i * 0.1 for i in range(1001)], {"X1": 100.0, "X2": 0.0}, "pseudo", "euler")
EXAMPLE_GBM = (["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"], list(np.arange(0.0, 10.0, 0.1)), {"X1": 1.0}, "pseudo", "runge-kutta")


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("name,model", [("example.py", EXAMPLE_PY), ("example.rs", EXAMPLE_RS), ("example_gbm.py", EXAMPLE_GBM)])
@pytest.mark.parametrize("scheme", ["as-written", "other"])
@pytest.mark.parametrize("rng_method", ["pseudo", "sobol"])
def test_reference_example_models(oracle, name, model, scheme, rng_method):
    eqs, times, init, _, sch = model
    if scheme == "other":
        sch = "euler" if sch == "runge-kutta" else "runge-kutta"
    N = 256
    U = oracle.Universe(eqs, times)
    assert (S.Universe(eqs, times).process_names, S.Universe(eqs, times).factor_names) == (U.names, U.factors)
    ref = oracle.simulate(U, init, N, sch, rng_method, seed=11)
    got = S.simulate(eqs, times, N, init, rng_method, sch, seed=11).to_numpy()
    assert got.shape == ref.shape == (N, len(times), len(eqs))
    # the payoff X2 = max(X1 - 100, 0) cancels X1 against 100 (X1 starts AT 100): its error is X1's absolute error, so it is
    # held to 1e-12 of X1's magnitude; every SDE process to 1e-12 relative
    levy = [i for i, e in enumerate(eqs) if e.lstrip().startswith("d")]
    payoff = [i for i in range(len(eqs)) if i not in levy]
    scale = float(np.abs(ref[:, :, levy]).max())

    def path_ok(i):
        return rel_err(got[i][:, levy], ref[i][:, levy]) <= 1e-12 and (not payoff or float(np.abs(got[i][:, payoff] - ref[i][:, payoff]).max()) <= 1e-12 * scale)

    ok = np.array([path_ok(i) for i in range(N)])
    flips = int((~ok).sum())
    assert flips <= (2 if any("dN" in e for e in eqs) else 0), (name, sch, rng_method, flips)
    if name == "example.rs":
        # dN1(0.5 * cos(t)): a negative intensity gives no jumps (increment.rs:183), a positive one does
        assert np.isfinite(got).all()


def test_reference_call_returns_the_reference_frame():
    # called with the reference's six arguments only, simulate returns the long frame the pyo3 function returns
    # (src/py_binding.rs:51-55): scenario:i32, time:f64, process_name:str, value:f64 in (scenario, time, process) order
    from sde_sim_rs.sde_sim_rs import simulate                # the compiled extension's module path (pyproject.toml:32)

    eqs, times, init, rng, sch = EXAMPLE_GBM
    df = simulate(processes_equations=eqs, time_steps=times, scenarios=50, initial_values=init, rng_method=rng, scheme=sch)
    assert list(df.columns) == ["scenario", "time", "process_name", "value"] and len(df) == 50 * len(times)
    assert str(df["scenario"].dtype).lower().startswith("int32") and str(df["value"].dtype).lower().startswith("float64")
    first = df.iloc[: len(times)] if hasattr(df, "iloc") else df.head(len(times))
    assert np.allclose(np.asarray(first["time"]), times) and float(np.asarray(first["value"])[0]) == 1.0
    dense = S.simulate(eqs, times, 50, init, rng, sch, seed=3)       # any extension keyword: the dense GPU tensor
    assert isinstance(dense, S.Filtration) and dense.shape == (50, len(times), 1)
    assert hasattr(S.simulate(eqs, times, 50, init, rng, sch, seed=3, frame=True), "columns")
    with pytest.raises(TypeError, match="unexpected keyword"):
        S.simulate(eqs, times, 50, init, rng, sch, sed=3)
