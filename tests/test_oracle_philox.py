"""Philox4x32-10 restatement (oracle/sde_oracle.cpp) against the known answers of Random123's kat_vectors
(Salmon et al., SC'11; the same generator as cuRAND's Philox4_32_10) and an independent pure-Python evaluation."""
import numpy as np


def _philox_py(ctr, key):
    c, k = list(ctr), list(key)
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & M, (p0 >> 32) ^ c[3] ^ k[1], p0 & M]
        k = [(k[0] + 0x9E3779B9) & M, (k[1] + 0xBB67AE85) & M]
    return c


KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers(oracle):
    for ctr, key, want in KATS:
        assert tuple(int(x) for x in oracle.philox4x32_10(ctr, key)) == want
        assert tuple(_philox_py(ctr, key)) == want


def test_philox_uniform_layout(oracle):
    seed, s = 0x0123456789ABCDEF, (1 << 33) + 12345
    for i in (0, 1, 2, 3, 4, 7, 1000):
        w = _philox_py((s & 0xFFFFFFFF, s >> 32, i >> 2, 0), (seed & 0xFFFFFFFF, seed >> 32))[i & 3]
        assert oracle.philox_uniform(seed, s, i) == (w + 0.5) * 2.0**-32
    U = oracle.Universe(["dA = ( 1.0 ) * dW1 + ( 1.0 ) * dW2"], [0.0, 1.0, 2.0, 3.0])
    u = oracle.uniforms(U, 3, "pseudo", seed=9, generator="philox", scenario_offset=5)
    assert u.shape == (3, 3, 2) and u[1, 2, 1] == oracle.philox_uniform(9, 6, 5)
    assert np.all((u > 0) & (u < 1))
