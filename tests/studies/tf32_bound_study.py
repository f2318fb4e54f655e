#!/usr/bin/env python
"""
CPU study for round 2: is a 3xTF32 tensor-core evaluation of  |q|^2 + |v|^2 - 2 q.v  (position / velocity dims)
plus the chord bound  2 - 2 cos(dh) <= dh^2  (heading) accurate enough to be used as a LOWER BOUND of the exact
float32 scan cost, and how many (query, node) pairs survive it?

TF32 is emulated by rounding float32 to 10 explicit mantissa bits (round-to-nearest-even on the bit pattern);
products of two TF32 numbers are exact in float32; accumulation is float32 (the tensor core's is at least as
accurate as truncating float32 adds, which is what `kappa` has to cover).  Features are centred on the sample
box so that |q|, |v| are as small as the problem allows.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from filter_study import grow          # noqa: E402


def tf32(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0xFFF + ((b >> 13) & 1)) & ~np.uint64(0x1FFF)
    return b.astype(np.uint32).view(np.float32)


def features(X, centre, heading=2):
    """[unwrapped dims centred..., cos h, sin h] in float32."""
    X = np.asarray(X, dtype=np.float64)
    cols = [d for d in range(X.shape[1]) if d != heading]
    F = np.concatenate(((X[:, cols] - centre[cols]), np.cos(X[:, [heading]]), np.sin(X[:, [heading]])), axis=1)
    return F.astype(np.float32)


def dot_3xtf32(A, B):
    """A (W x k), B (N x k) float32 -> W x N float32: hi*hi + hi*lo + lo*hi with float32 accumulation."""
    Ah = tf32(A); Al = tf32(A - Ah); Bh = tf32(B); Bl = tf32(B - Bh)
    out = np.zeros((A.shape[0], B.shape[0]), dtype=np.float32)
    for k in range(A.shape[1]):                       # one k at a time, float32 adds (worst-case ordering)
        for a, b in ((Al, Bh), (Ah, Bl), (Ah, Bh)):
            out = (out + (a[:, [k]] * b[None, :, k]).astype(np.float32)).astype(np.float32)
    return out


def exact_cost_f32(Q, V, heading=2):
    """The scan's own float32 chain (csrc/nearest.cuh), vectorised."""
    Q = Q.astype(np.float32); V = V.astype(np.float32)
    c = np.zeros((len(Q), len(V)), dtype=np.float32)
    for d in range(Q.shape[1]):
        if d == heading:
            two_pi = np.float32(6.283185307179586)
            rq = (Q[:, d] - two_pi * np.rint(Q[:, d] * np.float32(0.15915494309189535))).astype(np.float32)
            rv = (V[:, d] - two_pi * np.rint(V[:, d] * np.float32(0.15915494309189535))).astype(np.float32)
            e = np.abs(rq[:, None] - rv[None, :]).astype(np.float32)
            e = np.minimum(e, (two_pi - e).astype(np.float32))
        else:
            e = (Q[:, [d]] - V[None, :, d]).astype(np.float32)
        c = (e * e + c).astype(np.float32)          # fma would round once; this rounds twice -- irrelevant at the eps scale here
    return c


def main():
    nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    arena = 200.0 * np.sqrt(nodes / 100000.0)
    prob, orc = grow(nodes, arena)
    S = np.array(orc.state)
    Xq, _ = orc.sample_wave(99, 0, 0, 1024, prob.sample_space, prob.goal_bias, 10)
    ss = np.array(prob.sample_space)
    for label, scale in (("benchmark density, %.0f m arena" % arena, 1.0), ("same tree stretched to the 200 m arena", 200.0 / arena)):
        Sx = S.copy(); Qx = Xq.copy()
        Sx[:, :2] *= scale; Qx[:, :2] *= scale
        centre = np.zeros(S.shape[1]); centre[:2] = ss[:2].mean(axis=1) * scale
        FQ = features(Qx, centre); FV = features(Sx, centre)
        nq = (FQ.astype(np.float64) ** 2).sum(axis=1); nv = (FV.astype(np.float64) ** 2).sum(axis=1)
        dot = dot_3xtf32(FQ, FV).astype(np.float64)
        lb = nq[:, None] + nv[None, :] - 2.0 * dot
        # reference values in float64: the true bilinear bound and the exact cost
        lb64 = nq[:, None] + nv[None, :] - 2.0 * FQ.astype(np.float64).dot(FV.astype(np.float64).T)
        exact = exact_cost_f32(Qx, Sx).astype(np.float64)
        err = np.abs(lb - lb64)
        rel = err / (nq[:, None] + nv[None, :])
        best = exact.min(axis=1)
        print("== %s: |features|^2 up to %.0f" % (label, max(nq.max(), nv.max())))
        print("   3xTF32 bound vs float64 bilinear form: max abs err %.3g, max err/(|q|^2+|v|^2) = %.3g = 2^%.1f" % (
            err.max(), rel.max(), np.log2(rel.max())))
        for kappa_log2 in (-20, -19, -18):
            kappa = 2.0 ** kappa_log2
            safe = lb - kappa * (nq[:, None] + nv[None, :])
            viol = (safe > exact).sum()
            surv = (safe < best[:, None] * (1 + 1e-6)).sum(axis=1)
            print("   kappa 2^%d: bound violations %d of %d pairs; survivors per query: median %d, mean %.1f, max %d (of %d nodes)" % (
                kappa_log2, viol, exact.size, np.median(surv), surv.mean(), surv.max(), len(S)))


if __name__ == '__main__':
    main()
