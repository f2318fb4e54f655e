"""Philox4x32-10 restatement vs the Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from oracle.philox import philox4x32_10, uniforms

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_known_answers():
    for ctr, key, want in KAT:
        got = tuple(int(v) for v in philox4x32_10(*ctr, *key))
        assert got == want


def test_vectorised_matches_scalar():
    idx = np.arange(17, dtype=np.uint64)
    r = philox4x32_10(idx, 3, 1, 0, 42, 0)
    for i in range(17):
        s = philox4x32_10(i, 3, 1, 0, 42, 0)
        assert [int(v[i]) for v in r] == [int(v) for v in s]


def test_uniform_grid():
    u = uniforms(5, 0, np.arange(1000), 0, 7)
    assert u.shape == (1000, 7)
    assert (u >= 0).all() and (u < 1).all()
    assert np.array_equal(u * 16777216.0, np.round(u * 16777216.0))      # 2^-24 grid: exact in float32 too
    assert abs(u.mean() - 0.5) < 0.02
