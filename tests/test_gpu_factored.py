"""arithmetic="fast" lowers every Levy process in factored form (csrc/host/lower.cpp: factorise — coefficients flattened to
kappa * prod(factors), terms grouped by their state-dependent part, constants folded into per-step slots, common factors
taken out of the sum).  That is a small compiler pass: these models are built to walk its cases, and every one is held to the
oracle's strictly ordered f64 arithmetic (the reference's term-by-term order, src/sim/euler.rs:15-28 and
src/sim/runge_kutta.rs:45-97) on IDENTICAL draws, both schemes, uniform and non-uniform grids."""
import numpy as np
import pytest
import torch

from conftest import grid

import sde_sim_rs as S

pytestmark = pytest.mark.gpu

# (name, equations, initial values, Wiener mask of the factors in first-appearance order)
MODELS = [
    ("common-factor-two-groups",
     ["dX = ( 0.05 * X ) * dt + ( 0.2 * X * max(Y, 0.0)^0.5 ) * dW1 + ( -0.1 * max(Y, 0.0)^0.5 * X ) * dW2",
      "dY = ( 1.5 * (0.3 - Y) ) * dt + ( 0.25 * max(Y, 0.0)^0.5 ) * dW2"],
     {"X": 1.0, "Y": 0.3}, [True, True]),
    ("division-negation-e-pi",
     ["dX = ( -(X / 4.0) * pi() ) * dt + ( e() * 0.05 * -X ) * dW1 + ( 0.3 ) * dW2"],
     {"X": 2.0}, [True, True]),
    ("constant-and-zero-coefficients",
     ["dX = ( 0.1 ) * dt + ( 0.0 * X ) * dW1 + ( 2.0 ) * dW1 + ( 0.5 ) * dW2 + ( -0.25 ) * dW2"],
     {"X": -1.0}, [True, True]),
    ("time-factors",
     ["dX = ( sin(t) * X ) * dt + ( cos(t) * 0.2 ) * dW1 + ( t * 0.1 * X ) * dW1"],
     {"X": 1.5}, [True]),
    ("repeated-factors-and-comparisons",
     ["dX = ( 0.1 * X * X ) * dt + ( 0.05 * X^2 ) * dW1 + ( (X > 1.0) * 0.1 * X ) * dW1 + ( 0.02 * X * abs(X) * X ) * dW2"],
     {"X": 0.9}, [True, True]),
    ("three-processes-shared-factors",
     ["dA = ( 0.03 * A ) * dt + ( 0.2 * A ) * dW1",
      "dB = ( 0.01 * B * A ) * dt + ( 0.1 * B ) * dW1 + ( 0.15 * B ) * dW2",
      "dC = ( 0.5 * (A - C) ) * dt + ( 0.05 * (A + B) ) * dW2 + ( 0.02 * (A + B) ) * dW3"],
     {"A": 1.0, "B": 2.0, "C": 0.5}, [True, True, True]),
    ("levy-sees-algebraic",
     ["dX = ( 0.1 * (M - X) ) * dt + ( 0.2 * M ) * dW1",
      "M = 1.0 + 0.5 * X"],
     {"X": 1.0, "M": 1.5}, [True]),
    ("single-term-no-drift",
     ["dX = ( 0.3 * X / 2.0 ) * dW1"],
     {"X": 1.0}, [True]),
]
JUMP = ("jumps-with-literal-sizes",
        ["dX = ( 0.02 * X ) * dt + ( -0.3 * X ) * dN1(2.0 + abs(X)) + ( 0.1 * X ) * dW1 + ( 0.05 ) * dN1(2.0 + abs(X))"],
        {"X": 1.0}, [False, True])

NONUNIFORM = np.concatenate([[0.0], np.cumsum(np.linspace(2e-3, 6e-3, 61))])


def _inject(oracle, U, N, seed, wiener):
    u = oracle.uniforms(U, N, "pseudo", seed=seed)
    S_, K = u.shape[1], u.shape[2]
    buf = np.zeros((N, S_, K + 1))
    for k in range(K):
        buf[:, :, k] = oracle.icdf_normal(u[:, :, k].ravel()).reshape(N, S_) if wiener[k] else u[:, :, k]
    buf[:, :, K] = u[:, :, 0]
    return buf


def _both(oracle, eqs, times, init, N, wiener, scheme, seed=11):
    U = oracle.Universe(eqs, times)
    inj = _inject(oracle, U, N, seed, wiener)
    ref = oracle.simulate(U, init, N, scheme, inject=inj)
    dev = torch.from_numpy(inj).cuda()
    fast = S.Plan(S.Universe(eqs, times), scheme, "pseudo", inject=dev, arithmetic="fast").run(init, N).cpu().numpy()
    strict = S.Plan(S.Universe(eqs, times), scheme, "pseudo", inject=dev).run(init, N).cpu().numpy()
    return ref, strict, fast


@pytest.mark.parametrize("times", [grid(252, 64), NONUNIFORM], ids=["uniform", "nonuniform"])
@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
@pytest.mark.parametrize("name,eqs,init,wiener", MODELS, ids=[m[0] for m in MODELS])
def test_factored_lowering_matches_term_by_term_order(oracle, name, eqs, init, wiener, scheme, times):
    N = 300
    ref, strict, fast = _both(oracle, eqs, times, init, N, wiener, scheme)
    assert np.isfinite(ref).all()
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=(0, 1), keepdims=True) * 1e-3)   # values that cross zero: relative to the process scale
    assert np.max(np.abs(strict - ref) / scale) <= 1e-12
    assert np.max(np.abs(fast - ref) / scale) <= 1e-12, np.max(np.abs(fast - ref) / scale)


@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_factored_lowering_with_poisson_terms(oracle, scheme):
    name, eqs, init, wiener = JUMP
    N = 400
    ref, strict, fast = _both(oracle, eqs, grid(50, 40), init, N, wiener, scheme)
    # a jump count can flip where lambda*dt lands within an ulp of a CDF step: hold all but a few paths
    ok = np.max(np.abs(fast - ref) / np.maximum(np.abs(ref), 1e-3), axis=(1, 2)) <= 1e-12
    assert ok.mean() >= 0.99, ok.mean()
    assert (np.abs(ref[:, -1, 0] - init["X"]) > 0.05).mean() > 0.3        # the jumps are really there


@pytest.mark.parametrize("name,eqs,init,wiener", [MODELS[0], MODELS[5], MODELS[6]], ids=[MODELS[0][0], MODELS[5][0], MODELS[6][0]])
def test_factored_lowering_textbook_runge_kutta(oracle, name, eqs, init, wiener):
    """rk_variant="textbook" (k1 at the settled row, the separately named non-compat mode) through the same factored form."""
    times, N = grid(252, 48), 200
    U = oracle.Universe(eqs, times)
    inj = _inject(oracle, U, N, 13, wiener)
    ref = oracle.simulate(U, init, N, "runge-kutta", inject=inj, textbook_rk=True)
    dev = torch.from_numpy(inj).cuda()
    fast = S.Plan(S.Universe(eqs, times), "runge-kutta", "pseudo", inject=dev, arithmetic="fast", rk_variant="textbook").run(init, N).cpu().numpy()
    compat = S.Plan(S.Universe(eqs, times), "runge-kutta", "pseudo", inject=dev, arithmetic="fast").run(init, N).cpu().numpy()
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=(0, 1), keepdims=True) * 1e-3)
    assert np.max(np.abs(fast - ref) / scale) <= 1e-12
    assert np.max(np.abs(compat - ref) / scale) > 1e-9                    # the reference's stale-cache variant is a different scheme


def test_generated_source_shows_the_grouping():
    """What the pass did is readable in the plan's source (sde_plan_source): one root per stage, the two Wiener terms of X
    that share sqrt(Y+) X merged into one weight, X taken out of the drift + diffusion sum."""
    name, eqs, init, wiener = MODELS[0]
    src = S.Plan(S.Universe(eqs, grid(252, 8)), "euler", "sobol", scramble="xor", arithmetic="fast", icdf="fast").source
    step = src.split("sde_model_step(")[1]
    assert step.count("sde_f_sqrt_max0_fast(c[1])") == 2                 # one per process block (Euler)
    assert "fma(SDE_SLOT_2, zu[1], (SDE_SLOT_1 * zu[0]))" in step        # 0.2 sqrt(dt) z1 - 0.1 sqrt(dt) z2: one group
    assert "n0 = fma(c[0], fma(" in step                                  # X * (a + q * w) added onto row[0]


def _random_model(rng):
    """A random 2-process model whose coefficients are products / sums of bounded factors, so that paths stay finite whatever
    the draw: the shapes the factoring pass has to get right (shared and unshared factors, literals on either side, negations,
    division by literals, sums inside factors, repeated increments)."""
    names = ["X", "Y"]

    def factor():
        v = names[rng.integers(2)]
        w = names[rng.integers(2)]
        c = round(float(rng.uniform(0.2, 1.5)), 3)
        return [f"sin({v})", f"cos({c} * {w})", f"tanh({v} - {w})", f"max({v}, 0.0)^0.5", f"({c} - tanh({v}))", f"abs(sin({v} + t))",
                f"(sin({v})^2)", f"(1.0 + cos(t))", f"tanh({v})"][rng.integers(9)]

    def coeff():
        k = round(float(rng.uniform(-0.6, 0.6)), 3)
        parts = [factor() for _ in range(rng.integers(0, 4))]
        style = rng.integers(5)
        if not parts:
            return f"{k}"
        body = " * ".join(parts)
        if style == 0:
            return f"{k} * {body}"
        if style == 1:
            return f"{body} * {k}"
        if style == 2:
            return f"-({body}) * {abs(k)}"
        if style == 3:
            return f"{body} / {round(1.0 / max(abs(k), 0.05), 3)}"
        return f"{parts[0]} * {k} * " + " * ".join(parts[1:] + ["1.0"])

    eqs = []
    for n in names:
        terms = [f"( {coeff()} ) * dt"]
        for _ in range(rng.integers(1, 4)):
            terms.append(f"( {coeff()} ) * dW{rng.integers(1, 3)}")
        if rng.integers(3) == 0:
            terms.append(f"( {coeff()} ) * dt")
        eqs.append(f"d{n} = " + " + ".join(terms))
    return eqs


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("scheme", ["euler", "runge-kutta"])
def test_factored_lowering_random_models(oracle, seed, scheme):
    rng = np.random.default_rng(1000 + seed)
    eqs = _random_model(rng)
    times, init, N = grid(100, 40), {"X": 0.7, "Y": 0.4}, 128
    U = oracle.Universe(eqs, times)
    K = U.K
    if K == 0:
        pytest.skip("no stochastic factor drawn")
    ref, strict, fast = _both(oracle, eqs, times, init, N, [True] * K, scheme, seed=seed)
    assert np.isfinite(ref).all(), eqs
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=(0, 1), keepdims=True) * 1e-3)
    assert np.max(np.abs(strict - ref) / scale) <= 1e-12, eqs
    assert np.max(np.abs(fast - ref) / scale) <= 1e-11, (eqs, np.max(np.abs(fast - ref) / scale))
