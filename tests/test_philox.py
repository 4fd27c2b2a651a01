"""Philox4x32-10: Random123 known-answer vectors, and the three implementations
(oracle/philox.py, oracle/babyai_oracle.c, the kernels' Rng struct through the
host emulation) agree draw for draw."""
import numpy as np

import oracle as orc
from philox import PhiloxRandom, philox4x32_10

KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_known_answers_python():
    for ctr, key, out in KAT:
        assert philox4x32_10(ctr, key) == out


def test_known_answers_c_oracle():
    for ctr, key, out in KAT:
        assert orc.philox(ctr, key) == out


def test_randint_contract():
    r = PhiloxRandom(12345)
    assert r.randint(3, 4) == 3 and r.draws == 0          # a single-value range consumes nothing
    xs = [r.randint(0, 7) for _ in range(1000)]
    assert r.draws == 1000 and min(xs) == 0 and max(xs) == 6
    u = [r.uniform(0, 1) for _ in range(100)]
    assert all(0 <= v < 1 for v in u)


def test_stream_matches_c_oracle():
    # block b of the stream is philox((b, 0, 0, 0), (seed_lo, seed_hi))
    seed = 0x123456789ABCDEF0
    r = PhiloxRandom(seed)
    got = [r._u32() for _ in range(40)]
    exp = []
    for b in range(10):
        exp.extend(orc.philox((b, 0, 0, 0), (seed & 0xFFFFFFFF, seed >> 32)))
    assert got == exp
