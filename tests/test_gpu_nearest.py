"""
K1 parity (level 1): the device nearest-node search against the reference's
_costs_to_go + argmin / pruned argsort selection (planner.py:239-247, 340-350).
Index work is exact: the index must be identical, or (float32 only) a different index whose float64
cost is within 1e-6 relative of the true minimum.
"""
import numpy as np
import pytest

from conftest import DEMOS, load_golden, problem
from gpu_util import make_engine, nearest
from oracle.wave import WaveOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
@pytest.mark.parametrize('name', DEMOS)
def test_nearest_vs_reference_vectors(cuda_device, name, dtype):
    g = load_golden(name); prob = problem(name)
    eng = make_engine(prob, dtype)
    eng.load_nodes(g['state'], g['pID'])
    idx, cost = nearest(eng, g['hot_q'], use_ignore=False)
    assert np.array_equal(idx, g['hot_argmin'])
    assert np.allclose(cost, g['hot_mincost'], rtol=1e-12 if dtype == 'f64' else 1e-5)
    if len(g['hot_ignores']):
        eng.mark_ignore(g['hot_ignores'])
        idx2, _ = nearest(eng, g['hot_q'], use_ignore=True)
        assert np.array_equal(idx2, g['hot_pruned'])
        assert (idx2 != idx).any() or not np.isin(idx, g['hot_ignores']).any()


def _oracle_for(prob):
    return WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)


@pytest.mark.parametrize('dtype,N,W', [('f64', 3000, 300), ('f32', 20000, 1500), ('f32', 777, 5), ('f64', 1, 3),
                                       ('f32', 30001, 1), ('f64', 129, 257)])
def test_nearest_random_tree(cuda_device, dtype, N, W):
    """Random node cloud with headings spread over several turns (wrap matters), ragged sizes."""
    prob = problem('demo_boat_novice')
    rng = np.random.RandomState(N + W)
    ss = np.array(prob.sample_space)
    nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(N, 6)
    nodes[:, 2] = 15 * (rng.rand(N) - 0.5)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(W, 6)
    eng = make_engine(prob, dtype, node_capacity=N + 8, wave_capacity=W + 3)
    eng.load_nodes(nodes)
    idx, cost = nearest(eng, Q)
    orc = _oracle_for(prob); orc.load(nodes, [-1] * N, K=[None] * N)
    exact = 0
    for j in range(W):
        c = orc.costs(Q[j])
        best = int(np.argmin(c))
        if idx[j] == best:
            exact += 1
        else:
            assert dtype == 'f32' and c[idx[j]] <= c[best] * (1 + 1e-6) + 1e-9, (j, idx[j], best, c[idx[j]], c[best])
    assert exact >= W - max(1, W // 200)


def test_nearest_all_ignored_falls_back(cuda_device):
    """planner.py:240-245: if every node is ignored the plain argmin is used."""
    prob = problem('demo_boat_novice')
    rng = np.random.RandomState(3)
    nodes = 40 * rng.rand(50, 6)
    Q = 40 * rng.rand(20, 6)
    eng = make_engine(prob, 'f64')
    eng.load_nodes(nodes)
    base, _ = nearest(eng, Q, use_ignore=False)
    eng.mark_ignore(np.arange(50))
    idx, _ = nearest(eng, Q, use_ignore=True)
    assert np.array_equal(idx, base)
    eng.load_nodes(nodes)
    eng.mark_ignore(np.unique(base))                 # ignore exactly the winners: runner-ups must come back
    idx2, _ = nearest(eng, Q, use_ignore=True)
    assert not np.isin(idx2, np.unique(base)).any()
    orc = _oracle_for(prob); orc.load(nodes, [-1] * 50, K=[None] * 50)
    for k in np.unique(base):
        orc.ignore[int(k)] = True
    want, _ = orc.nearest(Q, True)
    assert np.array_equal(idx2, want)


def test_nearest_tie_takes_lowest_index(cuda_device):
    prob = problem('demo_boat_novice')
    nodes = np.zeros((600, 6)); nodes[:, 0] = np.arange(600) % 7      # many exact duplicates
    eng = make_engine(prob, 'f32')
    eng.load_nodes(nodes)
    Q = np.zeros((7, 6)); Q[:, 0] = np.arange(7)
    idx, cost = nearest(eng, Q)
    assert np.array_equal(idx, np.arange(7)) and np.all(cost == 0)


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_nearest_dense_metric(cuda_device, dtype):
    """Non-diagonal constant S (dense kernel path), S = lqr(x, 0)[0] of planner.py:348."""
    from lqrrt_b200 import models, configs
    base = problem('demo_boat_novice')
    rng = np.random.RandomState(11)
    A = rng.randn(6, 6); S = A.dot(A.T) + np.eye(6)
    system = models.System(base.system.dynamics, models.PDRotatedLqr(np.diag([120, 120, 350]), np.diag([120, 120, 100]), S=S),
                           models.Erf([2]), base.system.is_feasible)
    prob = configs.Problem('dense', system, base.x0, base.goal, base.goal_buffer, base.error_tol, base.sample_space,
                           base.goal_bias, base.horizon, base.dt, base.FPR)
    nodes = 40 * rng.rand(2500, 6); nodes[:, 2] = 12 * (rng.rand(2500) - 0.5)
    Q = 40 * rng.rand(200, 6)
    eng = make_engine(prob, dtype, node_capacity=4096, wave_capacity=256)
    eng.load_nodes(nodes)
    idx, cost = nearest(eng, Q)
    orc = _oracle_for(prob); orc.load(nodes, [-1] * len(nodes), K=[None] * len(nodes))
    for j in range(len(Q)):
        c = orc.costs(Q[j]); best = int(np.argmin(c))
        assert idx[j] == best or (dtype == 'f32' and c[idx[j]] <= c[best] * (1 + 1e-5))
        assert abs(cost[j] - c[idx[j]]) <= (1e-9 if dtype == 'f64' else 2e-3) * max(1.0, c[best])


def test_nearest_riccati_metric(cuda_device):
    """State-dependent S from the FD-Riccati provider (per-query metric kernel), parity unpinned by the
    reference: checked against the NumPy restatement."""
    prob = problem('c1')
    rng = np.random.RandomState(2)
    ss = np.array(prob.sample_space)
    nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(800, 2)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(64, 2)
    eng = make_engine(prob, 'f64')
    eng.load_nodes(nodes)
    idx, cost = nearest(eng, Q)
    orc = _oracle_for(prob); orc.load(nodes, [-1] * len(nodes), K=[None] * len(nodes))
    for j in range(len(Q)):
        c = orc.costs(Q[j]); best = int(np.argmin(c))
        assert idx[j] == best or c[idx[j]] <= c[best] * (1 + 1e-6)
        assert abs(cost[j] - c[idx[j]]) <= 1e-5 * max(1.0, c[best])


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_diagonal_non_identity_metric_vs_reference(cuda_device, dtype):
    """The ROS car's S = diag(1, 1, 1, 0, 0) (demos/lqrrt_ros/behaviors/car.py:65) through the pre-scaled form of
    the diagonal scan ((w q - w v)^2 per dim, w = sqrt(S_ii)) against the reference's own `_costs_to_go` with that
    S on the demo_car tree (tests/golden/make_golden_r2.py); also a non-representable weight set vs the oracle."""
    from lqrrt_b200 import configs, models
    g = load_golden('r2_extra'); base = problem('demo_car')
    car = base.system
    system = models.System(car.dynamics, models.PDRotatedLqr(car.lqr.kp, car.lqr.kd, S=g['ds_S']), models.Erf([2]),
                           car.is_feasible)
    prob = configs.Problem('ros_car_metric', system, base.x0, base.goal, base.goal_buffer, base.error_tol,
                           base.sample_space, base.goal_bias, base.horizon, base.dt, base.FPR)
    eng = make_engine(prob, dtype, node_capacity=len(g['ds_state']) + 8, wave_capacity=64)
    eng.load_nodes(g['ds_state'])
    idx, cost = nearest(eng, g['ds_q'])
    if dtype == 'f64':
        assert np.array_equal(idx, g['ds_argmin'])
        assert np.allclose(cost, g['ds_mincost'], rtol=1e-12, atol=1e-12)
    else:
        orc = _oracle_for(prob); orc.load(g['ds_state'], [-1] * len(g['ds_state']), K=[None] * len(g['ds_state']))
        for j in range(len(idx)):
            c = orc.costs(g['ds_q'][j])
            assert idx[j] == g['ds_argmin'][j] or c[idx[j]] <= c.min() * (1 + 1e-6) + 1e-9
        assert np.allclose(cost, g['ds_mincost'], rtol=1e-5, atol=1e-6)
    eng.close()
    # weights whose square roots are not exactly representable, boat (wrapped heading dim carries 3.3)
    bp = problem('demo_boat_novice')
    S = np.diag([2.0, 0.7, 3.3, 0.1, 5.0, 1.9])
    wsys = models.System(bp.system.dynamics, models.PDRotatedLqr(np.diag([120, 120, 350]), np.diag([120, 120, 100]), S=S),
                         models.Erf([2]), bp.system.is_feasible)
    wprob = configs.Problem('weighted', wsys, bp.x0, bp.goal, bp.goal_buffer, bp.error_tol, bp.sample_space,
                            bp.goal_bias, bp.horizon, bp.dt, bp.FPR)
    rng = np.random.RandomState(5)
    ss = np.array(bp.sample_space)
    nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(4000, 6); nodes[:, 2] = 15 * (rng.rand(4000) - 0.5)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(300, 6)
    eng = make_engine(wprob, dtype, node_capacity=4008, wave_capacity=304)
    eng.load_nodes(nodes)
    idx, cost = nearest(eng, Q)
    orc = _oracle_for(wprob); orc.load(nodes, [-1] * 4000, K=[None] * 4000)
    for j in range(len(Q)):
        c = orc.costs(Q[j])
        best = int(np.argmin(c))
        assert idx[j] == best or (dtype == 'f32' and c[idx[j]] <= c[best] * (1 + 1e-6) + 1e-9)
        assert abs(cost[j] - c[idx[j]]) <= (1e-11 if dtype == 'f64' else 2e-5) * max(1.0, c[best])
    eng.close()
