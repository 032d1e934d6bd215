"""Second, independent restatement of the reference schemes in pure Python — TEST INFRASTRUCTURE.

Purpose: cross-check oracle/sde_oracle.cpp (two implementations written separately from
the same reading of the Rust sources).  Small cases only.  Coefficients are Python
callables over a name->value dict, so this file shares no parser with the C++ oracle.
"parity unpinned": see sde_oracle.cpp.

Follows: src/filtration.rs:22-79 (rows + time-keyed cache), src/func.rs:32-42 (refresh
rule), src/sim/euler.rs:5-37, src/sim/runge_kutta.rs:5-107, src/proc/increment.rs:160-200,
src/rng/pseudo.rs:22-59.
"""
from __future__ import annotations

import math
import struct

MASK32 = 0xFFFFFFFF
MASK64 = 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------ ChaCha8 / seed_from_u64
def _rotl(x, r):
    return ((x << r) & MASK32) | (x >> (32 - r))


def chacha_block(key, counter, rounds, stream=0):
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *key,
         counter & MASK32, (counter >> 32) & MASK32, stream & MASK32, (stream >> 32) & MASK32]
    w = list(s)

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & MASK32; w[d] = _rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & MASK32; w[b] = _rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & MASK32; w[d] = _rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & MASK32; w[b] = _rotl(w[b] ^ w[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & MASK32 for a, b in zip(w, s)]


def seed_from_u64(state):
    key = []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & MASK64
        xs = (((state >> 18) ^ state) >> 27) & MASK32
        rot = state >> 59
        key.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & MASK32)
    return key


class ChaCha8:
    def __init__(self, seed):
        self.key = seed_from_u64(seed & MASK64)
        self.block = 0
        self.buf = []

    def next_u64(self):
        if not self.buf:
            self.buf = chacha_block(self.key, self.block, 8)
            self.block += 1
        lo, hi = self.buf[0], self.buf[1]
        self.buf = self.buf[2:]
        return lo | (hi << 32)

    def next_f64(self):
        return (self.next_u64() >> 11) * 2.0**-53


# ------------------------------------------------------------------ inverse CDFs
def icdf_normal(p):
    if p < 0.5:
        t = math.sqrt(-2.0 * math.log(p))
    else:
        t = math.sqrt(-2.0 * math.log(1.0 - p))
    c0, c1, c2 = 2.515517, 0.802853, 0.010328
    d1, d2, d3 = 1.432788, 0.189269, 0.001308
    x = t - ((c2 * t + c1) * t + c0) / (((d3 * t + d2) * t + d1) * t + 1.0)
    return -x if p < 0.5 else x


def icdf_poisson(u, lam):
    if lam <= 0.0:
        return 0
    p = math.exp(-lam)
    f = p
    k = 0
    while u > f and k < 200:
        k += 1
        p *= lam / k
        f += p
    return k


# ------------------------------------------------------------------ Sobol (Joe–Kuo, Gray code)
def sobol_direction_numbers(poly, minit):
    """poly/minit rows as in scipy's npz (row 0 = van der Corput)."""
    V = []
    for d in range(len(poly)):
        if d == 0:
            m = [1] * 64
        else:
            p = int(poly[d])
            s = p.bit_length() - 1
            m = [int(x) for x in minit[d][:s]]
            for i in range(s, 64):
                nv = m[i - s] ^ (m[i - s] << s)
                for k in range(1, s):
                    if (p >> (s - k)) & 1:
                        nv ^= m[i - k] << k
                m.append(nv)
        V.append([(m[i] << (63 - i)) & MASK64 for i in range(64)])
    return V


def sobol_point(V, n):
    g = n ^ (n >> 1)
    out = []
    for vd in V:
        x, b, gg = 0, 0, g
        while gg:
            if gg & 1:
                x ^= vd[b]
            gg >>= 1
            b += 1
        out.append(x)
    return out


# ------------------------------------------------------------------ model + schemes
class Levy:
    def __init__(self, name, terms):
        """terms: list of (coef_fn, kind, idx, lambda_fn) with kind in {'dt','dW','dN'}."""
        self.name, self.terms = name, terms


class Alg:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn


class Filtration:
    def __init__(self, procs, times, init):
        self.procs, self.times = procs, list(times)
        self.P = len(procs)
        self.raw = [[0.0] * self.P for _ in times]
        self.registry = {p.name: i for i, p in enumerate(procs)}
        self.time_registry = {struct.pack("<d", t): i for i, t in enumerate(times)}
        self.cache = {}
        self.cache_time = times[0]
        for k, v in init.items():
            if k in self.registry:
                self.raw[0][self.registry[k]] = v
        self.refresh(times[0])

    def refresh(self, time):
        self.cache_time = time
        self.cache["t"] = time
        ti = self.time_registry.get(struct.pack("<d", time), 0)
        for name, i in self.registry.items():
            self.cache[name] = self.raw[ti][i]

    def eval(self, fn, time):
        if time != self.cache_time:
            self.refresh(time)
        return fn(self.cache)


class StreamRng:
    """PseudoRng semantics (src/rng/pseudo.rs:35-59) over any f64 source."""

    def __init__(self, draw, K):
        self.draw, self.K, self.t, self.vals = draw, K, None, []

    def sample(self, t, k):
        if self.t != t:
            self.vals = [self.draw() for _ in range(self.K)]
            self.t = t
        return self.vals[k]


class TableRng:
    def __init__(self, values, K):
        self.values, self.K = values, K

    def sample(self, t, k):
        return self.values[t * self.K + k]


def _increment(term, F, rng, t, dts, sqrt_dts, z_inject=None):
    _, kind, idx, lam = term
    if kind == "dt":
        return dts[t]
    if kind == "dW":
        q = rng.sample(t, idx)
        return sqrt_dts[t] * (q if z_inject else icdf_normal(q))
    u = rng.sample(t, idx)
    return float(icdf_poisson(u, F.eval(lam, F.times[t]) * dts[t]))


def euler_step(F, rng, t, dts, sqrt_dts, z_inject=False):
    cur, nxt = F.times[t], F.times[t + 1]
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Levy):
            val = F.raw[t][p]
            for term in pr.terms:
                c = F.eval(term[0], cur)
                x = _increment(term, F, rng, t, dts, sqrt_dts, z_inject)
                val += c * x
            F.raw[t + 1][p] = val
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Alg):
            F.raw[t + 1][p] = F.eval(pr.fn, nxt)


def rk_step(F, rng, t, dts, sqrt_dts, z_inject=False, K=None):
    cur, nxt = F.times[t], F.times[t + 1]
    dt = nxt - cur
    sqrt_dt = math.sqrt(dt)
    u0 = rng.sample(t, K) if z_inject else rng.sample(t, 0)
    sk = 1.0 if u0 > 0.5 else -1.0
    incs = []
    for pr in F.procs:
        incs.append([_increment(term, F, rng, t, dts, sqrt_dts, z_inject) for term in pr.terms]
                    if isinstance(pr, Levy) else [])
    x_t = list(F.raw[t])
    k1 = [0.0] * F.P
    k2 = [0.0] * F.P
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Levy):
            for j, d in enumerate(incs[p]):
                k1[p] += F.eval(pr.terms[j][0], cur) * d
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Levy):
            pert = 0.0
            for term in pr.terms:
                if term[1] == "dW":
                    pert += F.eval(term[0], cur) * sk * sqrt_dt
            F.raw[t + 1][p] = x_t[p] + k1[p] + pert
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Levy):
            for j, d in enumerate(incs[p]):
                k2[p] += F.eval(pr.terms[j][0], nxt) * d
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Levy):
            F.raw[t + 1][p] = x_t[p] + 0.5 * (k1[p] + k2[p])
    for p, pr in enumerate(F.procs):
        if isinstance(pr, Alg):
            F.raw[t + 1][p] = F.eval(pr.fn, nxt)


def simulate_path(procs, times, init, rng, scheme, K, z_inject=False):
    F = Filtration(procs, times, init)
    dts = [times[i + 1] - times[i] for i in range(len(times) - 1)]
    sqrt_dts = [math.sqrt(d) for d in dts]
    for t in range(len(times) - 1):
        if scheme == "euler":
            euler_step(F, rng, t, dts, sqrt_dts, z_inject)
        elif scheme == "runge-kutta":
            rk_step(F, rng, t, dts, sqrt_dts, z_inject, K)
        else:
            raise NotImplementedError(scheme)
    return F.raw
