"""
CPU checks of the two floating-point facts the filtered nearest scans rest on (csrc/nearest.cuh
nearest_diag_prefix_kernel; csrc/grid.cuh relies on fact 1 as well):

1. the float32 cost chain c <- fma(e, e, c) never decreases, so the cost of the leading dims is a lower bound of
   the full cost in floating point, not only in real arithmetic;
2. |F(q)|^2 + |F(v)|^2 - 2 F(q).F(v) evaluated as 3xTF32 with float32 accumulation, minus kappa (|F(q)|^2 +
   |F(v)|^2) with kappa = 2^-19, never exceeds the exact float32 scan cost (chord bound for the heading).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'studies'))


def fma32(a, b, c):
    """float32 fma: a*b is exact in float64 (24 + 24 bits), one rounding of the sum to float64 then to float32
    (double rounding can differ from a true fma by one ulp in rare ties; irrelevant for monotonicity)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def test_cost_chain_is_monotone_in_float32():
    rng = np.random.RandomState(0)
    for scale in (1e-20, 1e-3, 1.0, 300.0, 1e18):
        e = (scale * rng.randn(200000, 12)).astype(np.float32)
        c = np.zeros(len(e), dtype=np.float32)
        for d in range(e.shape[1]):
            nxt = fma32(e[:, d], e[:, d], c)
            assert np.all(nxt >= c)
            c = nxt
    # adversarial: a huge accumulated cost and tiny increments must not round DOWN
    c = np.float32(1e12) * np.ones(1000, dtype=np.float32)
    e = (1e-3 * rng.rand(1000)).astype(np.float32)
    assert np.all(fma32(e, e, c) >= c)


def test_tf32_bound_never_exceeds_exact_cost():
    import tf32_bound_study as tb
    rng = np.random.RandomState(1)
    arena = 200.0
    lo = np.array([0, 0, -7, 0.9, -0.4, -0.2]); hi = np.array([arena, arena, 7, 1.1, 0.4, 0.2])
    V = lo + (hi - lo) * rng.rand(4000, 6)
    Q = lo + (hi - lo) * rng.rand(256, 6)
    V[:64] = Q[:64] + 1e-3 * rng.randn(64, 6)              # near-coincident pairs: the bound is tightest there
    centre = np.zeros(6); centre[:2] = arena / 2
    FQ = tb.features(Q, centre); FV = tb.features(V, centre)
    nq = (FQ.astype(np.float64) ** 2).sum(axis=1); nv = (FV.astype(np.float64) ** 2).sum(axis=1)
    lb = nq[:, None] + nv[None, :] - 2.0 * tb.dot_3xtf32(FQ, FV).astype(np.float64)
    exact = tb.exact_cost_f32(Q, V).astype(np.float64)
    safe = lb - 2.0 ** -19 * (nq[:, None] + nv[None, :])
    assert np.all(safe <= exact)
    # and it is a useful bound: within 0.2 of the exact cost wherever the heading difference is small
    dh = np.abs(np.arctan2(np.sin(Q[:, None, 2] - V[None, :, 2]), np.cos(Q[:, None, 2] - V[None, :, 2])))
    tight = dh < 0.3
    assert np.max((exact - safe)[tight]) < 0.2
    # chord inequality itself
    d = np.linspace(-np.pi, np.pi, 100001)
    assert np.all(2 - 2 * np.cos(d) <= d * d + 1e-15)
