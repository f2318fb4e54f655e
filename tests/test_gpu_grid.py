"""
The indexed nearest search (csrc/grid.cuh, LQ_OPT_NEAREST_GRID) against the brute-force scan it replaces: the
reference selection rule (planner.py:239-247, 340-350: smallest cost-to-go, lowest index among equal costs, nodes on
finished plans skipped unless nothing else is left) must come out bit for bit -- index AND cost -- whatever the
tree looks like: uniform, clustered far from the queries, outside the indexed extents, full of exact duplicates,
partly or wholly ignored; with extents taken from the sample box or from the loaded nodes, and after the box changed
under a loaded tree.  Planner level: the indexed search grows the very tree the scan grows, both wave lags.
"""
import numpy as np
import pytest

from conftest import problem
from gpu_util import make_engine, nearest

pytestmark = pytest.mark.gpu

OPT_NEAREST, OPT_GRID, OPT_GRID_MIN = 1, 9, 10


def _both(eng, Q, use_ignore=False, expect_grid=True):
    eng.set_option(OPT_GRID, 0)
    assert eng.get_option(OPT_GRID) == 0
    a = nearest(eng, Q, use_ignore)
    eng.set_option(OPT_GRID, 1); eng.set_option(OPT_GRID_MIN, 0)
    assert eng.get_option(OPT_GRID) == (1 if expect_grid else 0)
    b = nearest(eng, Q, use_ignore)
    return a, b


def _same(a, b):
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1].view(np.int64), b[1].view(np.int64))


def _cloud(prob, rng, count):
    ss = np.array(prob.sample_space)
    return ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(count, len(ss))


@pytest.mark.parametrize('dtype', ['f32', 'f64'])
@pytest.mark.parametrize('name,N,W,box', [('c4', 100000, 2048, 'before'), ('demo_boat_novice', 20000, 1500, 'before'),
                                           ('demo_boat_novice', 777, 5, 'none'), ('demo_boat_novice', 1, 3, 'before'),
                                           ('demo_boat_novice', 30001, 513, 'after'), ('demo_car', 9001, 700, 'none'),
                                           ('c5k', 5003, 600, 'before'), ('c3', 8, 512, 'after')])
def test_indexed_search_equals_scan(cuda_device, name, N, W, box, dtype):
    prob = problem(name)
    rng = np.random.RandomState(N + W)
    nodes = _cloud(prob, rng, N); Q = _cloud(prob, rng, W)
    if name != 'c5k':
        nodes[:, 2] = 15 * (rng.rand(N) - 0.5); Q[:, 2] = 15 * (rng.rand(W) - 0.5)      # headings over several turns
    # a fifth of the nodes and of the queries lie outside the sample box (rollouts are free to leave it)
    out = rng.rand(N) < 0.2; nodes[out, :2] += 60 * (rng.rand(out.sum(), 2) - 0.5)
    out = rng.rand(W) < 0.2; Q[out, :2] += 300 * (rng.rand(out.sum(), 2) - 0.5)
    eng = make_engine(prob, dtype, node_capacity=N + 8, wave_capacity=W + 3)
    if box == 'before':
        eng.set_sampling(prob.sample_space, prob.goal_bias, 10)
    eng.load_nodes(nodes)
    if box == 'after':                       # the extents move under a loaded tree: it is re-linked before the next search
        eng.set_sampling(prob.sample_space, prob.goal_bias, 10)
    a, b = _both(eng, Q)
    _same(a, b)
    if dtype == 'f64' and N <= 30001:        # and both are the reference's argmin
        e = Q[:40, None, :] - nodes[None, :, :]
        if name != 'c5k':
            e[:, :, 2] = np.arctan2(np.sin(e[:, :, 2]), np.cos(e[:, :, 2]))
        c = (e * e).sum(axis=2)
        assert np.array_equal(b[0][:40], c.argmin(axis=1))


@pytest.mark.parametrize('dtype', ['f32', 'f64'])
def test_indexed_search_clustered_tree_far_queries(cuda_device, dtype):
    """A young tree: every node within a few metres of the start, samples over the whole 200 m arena."""
    prob = problem('c4')
    rng = np.random.RandomState(2)
    N, W = 6000, 3000
    nodes = _cloud(prob, rng, N)
    nodes[:, :2] = 3.0 * rng.randn(N, 2) + 4.0
    nodes[:, 2] = 7 * (rng.rand(N) - 0.5)
    Q = _cloud(prob, rng, W)
    eng = make_engine(prob, dtype, node_capacity=N + 8, wave_capacity=W + 3)
    eng.set_sampling(prob.sample_space, prob.goal_bias, 10)
    eng.load_nodes(nodes)
    _same(*_both(eng, Q))


def test_indexed_search_ties_and_ignores(cuda_device):
    prob = problem('demo_boat_novice')
    rng = np.random.RandomState(5)
    nodes = np.zeros((5000, 6)); nodes[:, 0] = np.arange(5000) % 11; nodes[:, 1] = (np.arange(5000) // 11) % 3
    Q = np.zeros((600, 6)); Q[:, 0] = rng.randint(0, 11, 600); Q[:, 1] = rng.randint(0, 3, 600)
    eng = make_engine(prob, 'f32', node_capacity=5008, wave_capacity=640)
    eng.set_sampling(prob.sample_space, prob.goal_bias, 10)
    eng.load_nodes(nodes)
    a, b = _both(eng, Q)
    _same(a, b)
    assert np.array_equal(b[0], Q[:, 0].astype(int) + 11 * Q[:, 1].astype(int)) and np.all(b[1] == 0)
    eng.mark_ignore(np.unique(a[0]))                       # planner.py:240-245: fall to the same runner-ups
    c, d = _both(eng, Q, use_ignore=True)
    _same(c, d)
    assert not np.isin(d[0], np.unique(a[0])).any()
    eng.mark_ignore(np.arange(5000))                       # everything ignored: plain argmin again
    e, f = _both(eng, Q, use_ignore=True)
    _same(e, f)
    assert np.array_equal(f[0], a[0])


def test_indexed_search_not_applicable(cuda_device):
    """A wrapped leading dim (the pendulum's angle) cannot be indexed: the option stays off, results unchanged."""
    prob = problem('c1')
    rng = np.random.RandomState(9)
    nodes = _cloud(prob, rng, 3000); Q = _cloud(prob, rng, 300)
    eng = make_engine(prob, 'f32', node_capacity=3008, wave_capacity=320)
    eng.load_nodes(nodes)
    _same(*_both(eng, Q, expect_grid=False))


def _grow(prob, opts, lag, dtype='f32', nodes=20000, **kw):
    planner = prob.make_planner(dtype=dtype, seed=3, max_nodes=nodes, min_time=60, max_time=60, lag=lag, **kw)
    planner.engine_options = dict(opts)
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    tree = planner.tree
    return ok, tree.size, np.array(tree.state).tobytes(), list(tree.pID), planner.stats['waves'], planner.T


@pytest.mark.parametrize('lag', [1, 2])
@pytest.mark.parametrize('name', ['c3', 'c4', 'demo_car'])
def test_indexed_search_grows_the_same_tree(cuda_device, name, lag):
    prob = problem(name)
    base = _grow(prob, {OPT_GRID: 0}, lag)
    got = _grow(prob, {OPT_GRID: 1, OPT_GRID_MIN: 0}, lag)
    assert got == base
    auto = _grow(prob, {}, lag)                               # the default (indexed above its size threshold)
    assert auto == base


def test_indexed_search_with_goal_paths(cuda_device):
    """Plans that reach the goal early: ignore marks (both copies with pipelined waves) and the index together."""
    prob = problem('demo_boat_novice')
    for lag in (1, 2):
        base = _grow(prob, {OPT_GRID: 0}, lag, dtype='f64', nodes=4000, wave=64)
        got = _grow(prob, {OPT_GRID: 1, OPT_GRID_MIN: 0}, lag, dtype='f64', nodes=4000, wave=64)
        assert got == base
