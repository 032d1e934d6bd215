"""pytest configuration: markers, import paths, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sde-sim-rs_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun / the driver's GPU tier)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.build()
    return orc


GBM_EQ = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
HESTON_EQ = [
    "dS = ( 0.05 * S ) * dt + ( max(v, 0.0)^0.5 * S ) * dW1",
    "dv = ( 2.0 * (0.04 - v) ) * dt + ( -0.21 * max(v, 0.0)^0.5 ) * dW1 + ( 0.2142428528562855 * max(v, 0.0)^0.5 ) * dW2",
]


def grid(D, S=None):
    S = D if S is None else S
    return [k / D for k in range(S + 1)]


def basket_equations(n_assets=64, rho=0.5):
    """C4: correlated GBM basket written the way the reference expresses correlation —
    shared dW names with explicit Cholesky loadings (src/proc/util.rs:145-146)."""
    import numpy as np

    corr = np.full((n_assets, n_assets), rho)
    np.fill_diagonal(corr, 1.0)
    L = np.linalg.cholesky(corr)
    eqs = []
    for i in range(n_assets):
        sig = 0.1 + 0.2 * i / max(n_assets - 1, 1)
        terms = [f"( 0.05 * S{i} ) * dt"] + [f"( {sig * L[i, j]:.17g} * S{i} ) * dW{j + 1}" for j in range(i + 1)]
        eqs.append(f"dS{i} = " + " + ".join(terms))
    init = {f"S{i}": 100.0 for i in range(n_assets)}
    return eqs, init
