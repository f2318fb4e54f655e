"""
Engine options (include/lqrrt_b200.h lq_set_option) change how the work is issued, never the result:

* LQ_OPT_NEAREST_KERNEL = packed: the float32 diagonal-metric scan on the FP32x2 pipe with a two-level
  argmin must return bit-identical (index, cost) to the scalar scan -- same lowest-index tie-break
  (planner.py:247 np.argmin), same ignore handling (planner.py:240-245), ragged sizes included;
* LQ_OPT_SAMPLE_PREFETCH: the sampler on a side stream must grow the same tree, bit for bit.
"""
import numpy as np
import pytest

from conftest import problem
from gpu_util import make_engine, nearest

pytestmark = pytest.mark.gpu

OPT_NEAREST, OPT_PREFETCH, OPT_CTAS = 1, 2, 3
SCALAR, PACKED, PACKED_Q8, TWO_LEVEL = 0, 1, 2, 3


def _both(eng, Q, use_ignore=False, kernel=PACKED, ctas=0):
    eng.set_tuning(0, 4)                         # large-wave launch shape (several queries per thread) for any W
    eng.set_option(OPT_NEAREST, SCALAR); eng.set_option(OPT_CTAS, 8)
    a = nearest(eng, Q, use_ignore)
    eng.set_option(OPT_NEAREST, kernel); eng.set_option(OPT_CTAS, ctas)
    assert eng.get_option(OPT_NEAREST) == kernel
    b = nearest(eng, Q, use_ignore)
    return a, b


def _cloud(prob, rng, count, turns=15.0):
    ss = np.array(prob.sample_space)
    pts = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(count, len(ss))
    return pts


@pytest.mark.parametrize('name,N,W', [('demo_boat_novice', 20000, 1500), ('demo_boat_novice', 777, 5),
                                      ('demo_boat_novice', 1, 3), ('demo_boat_novice', 30001, 513),
                                      ('demo_boat_novice', 129, 1027), ('demo_boat_novice', 8, 512),
                                      ('demo_car', 9001, 700), ('c5k', 5003, 600), ('c4', 100000, 2048)])
@pytest.mark.parametrize('kernel,ctas', [(PACKED, 0), (PACKED_Q8, 8), (TWO_LEVEL, 16), (PACKED, 3)])
def test_packed_scan_is_bit_identical(cuda_device, name, N, W, kernel, ctas):
    prob = problem(name)
    rng = np.random.RandomState(N + W)
    nodes = _cloud(prob, rng, N); Q = _cloud(prob, rng, W)
    if name != 'c5k':
        nodes[:, 2] = 15 * (rng.rand(N) - 0.5)           # headings over several turns: the wrap matters
        Q[:, 2] = 15 * (rng.rand(W) - 0.5)
    eng = make_engine(prob, 'f32', node_capacity=N + 8, wave_capacity=W + 3)
    eng.load_nodes(nodes)
    (i0, c0), (i1, c1) = _both(eng, Q, kernel=kernel, ctas=ctas)
    assert np.array_equal(i0, i1)
    assert np.array_equal(c0.view(np.int64), c1.view(np.int64))


def test_packed_scan_ties_and_ignores(cuda_device):
    prob = problem('demo_boat_novice')
    rng = np.random.RandomState(5)
    # many exact duplicates, spread over sub-blocks, tiles and splits: the lowest index must win
    nodes = np.zeros((5000, 6)); nodes[:, 0] = np.arange(5000) % 11; nodes[:, 1] = (np.arange(5000) // 11) % 3
    Q = np.zeros((600, 6)); Q[:, 0] = rng.randint(0, 11, 600); Q[:, 1] = rng.randint(0, 3, 600)
    eng = make_engine(prob, 'f32', node_capacity=5008, wave_capacity=640)
    eng.load_nodes(nodes)
    (i0, c0), (i1, c1) = _both(eng, Q)
    assert np.array_equal(i0, i1) and np.all(c1 == 0)
    assert np.array_equal(i1, Q[:, 0].astype(int) + 11 * Q[:, 1].astype(int))
    # ignore the winners (planner.py:240-245): both scans must fall to the same runner-ups
    eng.mark_ignore(np.unique(i0))
    (j0, d0), (j1, d1) = _both(eng, Q, use_ignore=True)
    assert np.array_equal(j0, j1) and np.array_equal(d0.view(np.int64), d1.view(np.int64))
    assert not np.isin(j1, np.unique(i0)).any()
    # everything ignored: plain argmin again
    eng.mark_ignore(np.arange(5000))
    (k0, _), (k1, _) = _both(eng, Q, use_ignore=True)
    assert np.array_equal(k0, i0) and np.array_equal(k1, i0)


def _grow(prob, **kw):
    planner = prob.make_planner(dtype='f32', seed=3, max_nodes=20000, min_time=60, max_time=60, **kw)
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    tree = planner.tree
    return ok, tree.size, np.array(tree.state), list(tree.pID), planner.stats['waves'], planner.T


@pytest.mark.parametrize('name', ['c3', 'c4'])
def test_options_grow_the_same_tree(cuda_device, name):
    prob = problem(name)
    base = _grow(prob, nearest_kernel='scalar', prefetch=False)
    for kw in (dict(nearest_kernel='packed', prefetch=False), dict(nearest_kernel='scalar', prefetch=True),
               dict(nearest_kernel='packed', prefetch=True)):
        got = _grow(prob, **kw)
        assert got[0] == base[0] and got[1] == base[1] and got[4] == base[4], kw
        assert got[3] == base[3], kw
        assert np.array_equal(got[2].view(np.int64), base[2].view(np.int64)), kw
        assert got[5] == base[5]


def test_prefetch_across_plans_and_wave_sizes(cuda_device):
    """Replanning pattern (lqrrt_node.py:389-500): the double-buffered sampler must survive resets, odd wave
    counts and fixed small waves."""
    prob = problem('c3')
    out = []
    for pf in (False, True):
        planner = prob.make_planner(dtype='f32', seed=11, max_nodes=3000, min_time=60, max_time=60, prefetch=pf,
                                    wave=96)
        sizes = []
        for rep in range(3):
            planner.update_plan(prob.x0 + 0.01 * rep, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
            sizes.append((planner.tree.size, planner.stats['waves'], np.array(planner.tree.state).tobytes()))
        out.append(sizes)
    assert out[0] == out[1]


# ---- the partial-distance scan variant and non-identity weights (first run on a B200 in round 2)
PREFIX = 4


def _weighted_problem():
    """Diagonal, non-identity S (weights sqrt(S_ii) that are not exactly representable products)."""
    from lqrrt_b200 import models, configs
    base = problem('demo_boat_novice')
    S = np.diag([2.0, 0.7, 3.3, 0.1, 5.0, 1.9])
    system = models.System(base.system.dynamics, models.PDRotatedLqr(np.diag([120, 120, 350]), np.diag([120, 120, 100]), S=S),
                           models.Erf([2]), base.system.is_feasible)
    return configs.Problem('weighted', system, base.x0, base.goal, base.goal_buffer, base.error_tol, base.sample_space,
                           base.goal_bias, base.horizon, base.dt, base.FPR)


@pytest.mark.parametrize('name,N,W', [('demo_boat_novice', 20000, 1500), ('demo_boat_novice', 129, 1027), ('c4', 100000, 2048),
                                      ('c5k', 5003, 600), ('demo_car', 9001, 700)])
def test_prefix_filter_scan_is_bit_identical(cuda_device, name, N, W):
    prob = problem(name)
    rng = np.random.RandomState(N + W)
    nodes = _cloud(prob, rng, N); Q = _cloud(prob, rng, W)
    if name != 'c5k':
        nodes[:, 2] = 15 * (rng.rand(N) - 0.5); Q[:, 2] = 15 * (rng.rand(W) - 0.5)
    eng = make_engine(prob, 'f32', node_capacity=N + 8, wave_capacity=W + 3)
    eng.load_nodes(nodes)
    (i0, c0), (i1, c1) = _both(eng, Q, kernel=PREFIX)
    assert np.array_equal(i0, i1)
    assert np.array_equal(c0.view(np.int64), c1.view(np.int64))
    eng.mark_ignore(np.unique(i0)[::2])
    (j0, d0), (j1, d1) = _both(eng, Q, use_ignore=True, kernel=PREFIX)
    assert np.array_equal(j0, j1) and np.array_equal(d0.view(np.int64), d1.view(np.int64))


@pytest.mark.parametrize('kernel', [PACKED, PACKED_Q8, TWO_LEVEL, PREFIX])
def test_variants_agree_for_non_identity_weights(cuda_device, kernel):
    """Pre-scaling is a rounded multiply of its own in every variant (csrc/nearest.cuh prescale), so the
    variants agree bit for bit for any diagonal S, not only S = I; small and large launch shapes too."""
    prob = _weighted_problem()
    rng = np.random.RandomState(17)
    nodes = _cloud(prob, rng, 30000); Q = _cloud(prob, rng, 900)
    nodes[:, 2] = 15 * (rng.rand(len(nodes)) - 0.5); Q[:, 2] = 15 * (rng.rand(len(Q)) - 0.5)
    eng = make_engine(prob, 'f32', node_capacity=30008, wave_capacity=904)
    eng.load_nodes(nodes)
    (i0, c0), (i1, c1) = _both(eng, Q, kernel=kernel)
    assert np.array_equal(i0, i1) and np.array_equal(c0.view(np.int64), c1.view(np.int64))
    eng.set_tuning(0, 1); eng.set_option(OPT_NEAREST, SCALAR)          # one query per thread (small-wave shape)
    i2, c2 = nearest(eng, Q)
    assert np.array_equal(i0, i2) and np.array_equal(c0.view(np.int64), c2.view(np.int64))
