"""Known-answer tests of the random stream the oracle and the CUDA path share.

Philox4x32-10 vectors are the published Random123 known-answer tests (Salmon et al., SC'11, kat_vectors:
`philox4x32 10`).  The stream itself replaces R's global RNG, which bcf draws from (SURVEY 8b "RNG")."""
import numpy as np
import pytest

from oracle import orc

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,expect", KAT)
def test_philox4x32_10_known_answers(ctr, key, expect):
    assert tuple(int(v) for v in orc.philox4x32_10(ctr, key)) == expect


def test_uniforms_are_53_bit_open_interval():
    us = np.array([orc.rng_probe(42, 0, s, t, 0, 0)[:2] for s in range(40) for t in range(25)])
    assert us.min() > 0.0 and us.max() < 1.0
    # (k + 0.5) / 2^53: scaled values are half-integers
    k = us * 2.0 ** 53 - 0.5
    assert np.all(k == np.floor(k))
    assert abs(us.mean() - 0.5) < 0.02


def test_stream_is_counter_addressed():
    a = orc.rng_probe(7, 1, 10, 3, 3, 99)
    assert np.array_equal(a, orc.rng_probe(7, 1, 10, 3, 3, 99))
    for other in ((8, 1, 10, 3, 3, 99), (7, 2, 10, 3, 3, 99), (7, 1, 11, 3, 3, 99), (7, 1, 10, 4, 3, 99),
                  (7, 1, 10, 3, 2, 99), (7, 1, 10, 3, 3, 98)):
        assert not np.array_equal(a, orc.rng_probe(*other))


def test_normal_and_gamma_moments():
    z = np.array([orc.rng_probe(1, 0, s, 0, 3, i)[2] for s in range(200) for i in range(50)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    for a in (0.4, 1.0, 2.5, 250000.0):
        g = np.array([orc.rng_gamma(3, 0, s, 15, a) for s in range(4000)])
        assert abs(g.mean() / a - 1.0) < 4.0 / np.sqrt(a * 4000) + 1e-3
        assert abs(g.var() / a - 1.0) < 0.15
