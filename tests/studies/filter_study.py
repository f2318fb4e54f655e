#!/usr/bin/env python
"""
CPU study for round 2 (no GPU needed; lives under tests/ because it drives the wave ORACLE, which only test
infrastructure may import): how often would a partial-distance filter in the nearest scan have
to fall back to the full evaluation?

The exact scan accumulates  c = sum_d (w_d (q_d - v_d))^2  in dim order; every prefix of that sum is a
rigorous lower bound of the full cost in floating point (all terms >= 0, rounding is monotone).  A scan that
evaluates only the first P dims (the positions) and the remaining dims only for sub-blocks of 8 nodes in which
some query of the warp (32 lanes x 4 queries) has prefix < its running best stays bit-identical to the exact
scan.  Its pay-off depends on (a) how tight the running best is -- seeded from a coarse (x, y) grid holding
the latest K nodes per cell -- and (b) node density.  This script grows a C4-like tree with the wave oracle
at the benchmark's node density (same nodes per square metre as 100k nodes on the 200 m arena) and reports the
trigger probability per (warp, sub-block).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lqrrt_b200 import configs          # noqa: E402
from oracle.wave import WaveOracle      # noqa: E402


def grow(target_nodes, arena, seed=0):
    prob = configs.c4_boat_advanced(seed, count=max(1, int(500 * (arena / 200.0) ** 2)), arena=arena, max_nodes=target_nodes)
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)
    orc.reset(prob.x0)
    wave = 0
    t0 = time.time()
    while orc.size < target_nodes:
        W = int(min(2048, max(32, 2 * orc.size)))
        X, _ = orc.sample_wave(7, wave, 0, W, prob.sample_space, prob.goal_bias, 10)
        orc.expand_wave(X)
        wave += 1
        print("wave %d: %d nodes (%.0f s)" % (wave, orc.size, time.time() - t0), flush=True)
    return prob, orc


def study(prob, orc, n_queries=2048, cell=1.5, keep=4, prefix=2, sb=8, warp_queries=128):
    S = np.array(orc.state[:, :], dtype=np.float64)
    N = len(S)
    Xq, _ = orc.sample_wave(99, 0, 0, n_queries, prob.sample_space, prob.goal_bias, 10)

    def cost_full(q, V):
        d = q[None, :] - V
        d[:, 2] = np.arctan2(np.sin(d[:, 2]), np.cos(d[:, 2]))
        return (d * d).sum(axis=1)
    # grid of the latest `keep` nodes per (x, y) cell
    lo = np.array([s[0] for s in prob.sample_space[:2]]); hi = np.array([s[1] for s in prob.sample_space[:2]])
    g = np.maximum(1, np.ceil((hi - lo) / cell).astype(int))
    cells = {}
    ij = np.clip(((S[:, :2] - lo) / cell).astype(int), 0, g - 1)
    for node in range(N):
        cells.setdefault((ij[node, 0], ij[node, 1]), []).append(node)
    full_pass = {0.1: 0, 0.5: 0, 2.0: 0}
    n_sub = (N + sb - 1) // sb
    pre = np.zeros((n_queries, N))
    best = np.zeros(n_queries); ub = np.zeros(n_queries)
    for j, q in enumerate(Xq):
        c = cost_full(q, S.copy())
        best[j] = c.min()
        # a bound on ALL dims (tensor-core dot-product form): the heading term is relaxed to the chord
        # 2 - 2 cos(dh) <= dh^2 and the result is trusted only to an absolute eps
        d = q[None, :] - S
        lb = (d[:, [0, 1, 3, 4, 5]] ** 2).sum(axis=1) + 2.0 - 2.0 * np.cos(d[:, 2])
        for eps in full_pass:
            full_pass[eps] += int((lb - eps < best[j]).sum())
        qi = np.clip(((q[:2] - lo) / cell).astype(int), 0, g - 1)
        cand = []
        for a in (-1, 0, 1):
            for b in (-1, 0, 1):
                cand += cells.get((qi[0] + a, qi[1] + b), [])[-keep:]
        ub[j] = c[cand].min() if cand else np.inf
        dxy = q[None, :prefix] - S[:, :prefix]
        pre[j] = (dxy * dxy).sum(axis=1)
    # per (warp of `warp_queries` queries, sub-block of `sb` nodes): does any query need the full evaluation?
    pad = n_sub * sb - N
    prep = np.pad(pre, ((0, 0), (0, pad)), constant_values=np.inf).reshape(n_queries, n_sub, sb).min(axis=2)
    for bound, name in ((ub, 'grid-seeded bound'), (best, 'oracle bound (true minimum)')):
        need = prep < bound[:, None]                      # (query, sub-block)
        warps = need.reshape(n_queries // warp_queries, warp_queries, n_sub).any(axis=1)
        print("%-28s per-query sub-block rate %.4f   per-warp trigger rate %.3f" % (name, need.mean(), warps.mean()))
    for eps, cnt in sorted(full_pass.items()):
        print("all-dims bound, eps %.1f: %.2f candidate nodes per query (pass rate %.2e)" % (eps, cnt / n_queries, cnt / n_queries / N))
    fin = np.isfinite(ub)
    print("nodes %d  density %.2f /m^2  queries %d  median best cost %.3f  median grid bound %.3f  (bound/best median %.2f, no-bound %.1f %%)" % (
        N, N / np.prod(hi - lo), n_queries, np.median(best), np.median(ub[fin]), np.median(ub[fin] / np.maximum(best[fin], 1e-9)), 100 * (1 - fin.mean())))


if __name__ == '__main__':
    nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    arena = 200.0 * np.sqrt(nodes / 100000.0)
    prob, orc = grow(nodes, arena)
    for cell in (1.0, 2.0):
        for keep in (1, 4):
            print("cell %.1f m, keep %d" % (cell, keep))
            study(prob, orc, cell=cell, keep=keep)
