"""
ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/models.py header).

Philox4x32-10 (Salmon et al., SC'11: "Parallel random numbers: as easy as
1, 2, 3"), restated in NumPy uint64 arithmetic; the device twin is
csrc/sample.cuh.  The reference samples from NumPy's global MT19937
(planner.py:204-205), which is inherently serial; the engine's throughput mode
replaces it with this counter-based generator keyed by (seed, wave, sample,
try) so that any partition of a wave over GPUs draws the same numbers.
Known-answer vectors from the Random123 distribution are in
tests/test_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint64(0x9E3779B9)
W1 = np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)
SH = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Counters/keys: arrays (or scalars) of uint32 values held in uint64.
    Returns four uint64 arrays holding uint32 outputs."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0) & MASK
    k1 = np.uint64(k1) & MASK
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0 = p0 >> SH; lo0 = p0 & MASK
        hi1 = p1 >> SH; lo1 = p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def uniforms(seed, wave, sample, attempt, count):
    """`count` uniforms in [0,1) on a 2^-24 grid for each element of `sample`:
    block b supplies uniforms 4b..4b+3 from philox(counter=(sample, wave,
    attempt, b), key=(seed_lo, seed_hi)); u = (r >> 8) * 2^-24."""
    sample = np.atleast_1d(np.asarray(sample, dtype=np.uint64))
    out = np.zeros((len(sample), 4 * ((count + 3) // 4)))
    for b in range((count + 3) // 4):
        r = philox4x32_10(sample, np.uint64(wave), np.uint64(attempt), np.uint64(b),
                          np.uint64(seed) & MASK, (np.uint64(seed) >> SH) & MASK)
        for j in range(4):
            out[:, 4 * b + j] = (r[j] >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0)
    return out[:, :count]
