"""
Planner-level parity (level 2) and API behaviour through the public classes.

wave=1, rng='numpy', dtype='f64' reproduces the reference iteration for iteration: same global
MT19937 stream, same fake clock => same tree as tests/golden (the unmodified reference's output).
Tolerance: parents and edge lengths identical, states within 1e-9, plan duration T identical
(the 1 % path-cost bar of BASELINE.json is met with 0 %).
"""
import numpy as np
import pytest

from conftest import DEMOS, load_golden, problem, setup_draws

pytestmark = pytest.mark.gpu


class WaveClock(object):
    """Fake sys_time: ticks once per call; update_plan calls it once before the loop and once per wave."""

    def __init__(self):
        self.t = -1

    def __call__(self):
        self.t += 1
        return self.t


@pytest.mark.parametrize('name', DEMOS)
def test_wave1_reproduces_reference_tree(cuda_device, name):
    g = load_golden(name); prob = problem(name)
    K = int(g['K'])
    planner = prob.make_planner(dtype='f64', wave=1, rng='numpy', min_time=K, max_time=K, sys_time=WaveClock())
    np.random.seed(0)
    np.random.rand(setup_draws(name))
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    assert ok
    assert planner.stats['waves'] == K
    tree = planner.tree
    assert tree.size == len(g['state'])
    assert tree.pID == g['pID'].tolist()
    err = np.abs(tree.state - g['state']).max(axis=1)
    # The car-like boats' dynamics are discontinuous at vx == 0 (demo_boat_advanced.py:122-128: the yaw
    # rate is rescaled for vx > 0 or vx < 0 but left alone at exactly 0, right after vx was clamped to 0),
    # so a last-bit difference in the thruster mix can flip that branch on a near-stationary edge.  Every
    # other node must agree to 1e-9; the few knife-edge ones stay within 1e-4 (same parents, same plan).
    assert np.mean(err <= 1e-9) >= 0.95 and err.max() <= 1e-4, (np.mean(err <= 1e-9), err.max())
    if name in ('demo_boat_novice', 'demo_car'):
        assert err.max() <= 1e-9
    assert [len(e) for e in tree.x_seq] == g['edge_len'].tolist()
    assert planner.plan_reached_goal == bool(g['reached'])
    assert planner.T == pytest.approx(float(g['T']), rel=0, abs=1e-12)
    assert planner.node_seq == g['node_seq'].tolist()
    assert np.abs(np.array(planner.x_seq) - g['x_seq']).max() <= 1e-7
    assert np.abs(np.array(planner.u_seq) - g['u_seq']).max() <= 1e-4
    # interpolators (planner.py:451-464)
    tq = 0.37 * planner.T
    i = int(tq / planner.dt)
    assert np.allclose(planner.get_state(tq), g['x_seq'][i] + (g['x_seq'][i + 1] - g['x_seq'][i]) * (tq / planner.dt - i), atol=1e-6)
    assert np.allclose(planner.get_state(1e9), g['x_seq'][-1], atol=1e-6)


@pytest.mark.parametrize('name', ['demo_boat_novice', 'demo_car'])
def test_f32_wave1_path_cost_within_one_percent(cuda_device, name):
    """BASELINE.json: final path cost within 1 % of the reference on a fixed seed, float32 engine."""
    g = load_golden(name); prob = problem(name)
    K = int(g['K'])
    planner = prob.make_planner(dtype='f32', wave=1, rng='numpy', min_time=K, max_time=K, sys_time=WaveClock())
    np.random.seed(0)
    np.random.rand(setup_draws(name))
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    assert planner.plan_reached_goal == bool(g['reached'])
    assert abs(planner.T - float(g['T'])) <= 0.01 * float(g['T'])
    m = min(planner.tree.size, len(g['state']))
    same = planner.tree.pID[:m] == g['pID'].tolist()[:m]
    if same and planner.tree.size == len(g['state']):
        assert np.abs(planner.tree.state - g['state']).max() <= 1e-3     # fp64 -> fp32 tolerance on node states


def test_throughput_mode_plans_and_reads_back(cuda_device):
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f32', seed=3, min_time=0.0, max_time=20.0, max_nodes=60000, wave_max=512)
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, finish_on_goal=True)
    assert ok and planner.plan_reached_goal
    tree = planner.tree
    assert tree.size > 100 and np.isfinite(planner.T)
    # the plan is a chain seed -> goal region, each edge ends in its node, consecutive states move < 1 step of top speed
    assert planner.node_seq[0] == 0
    for a, b in zip(planner.node_seq[:-1], planner.node_seq[1:]):
        assert tree.pID[b] == a
    xs = np.array(planner.x_seq)
    assert np.allclose(xs[0], prob.x0)
    assert np.all(np.linalg.norm(np.diff(xs[:, :2], axis=0), axis=1) < 0.2)
    assert len(planner.t_seq) == len(planner.x_seq) == len(planner.u_seq)
    # planner.py:300-303: finish_on_goal extends x_seq/t_seq but leaves T as the plan's own duration
    assert planner.T <= len(planner.x_seq) * planner.dt + 1e-9
    # planner.py:299: the extra node's state is the exact goal, its edge is the force-arrive rollout, and that
    # rollout is what the plan was extended by
    last = planner.node_seq[-1]
    if np.array_equal(tree.state[last], prob.goal):
        assert tree.pID[last] == planner.node_seq[-2]
        edge = np.array(tree.x_seq[last])
        assert len(edge) >= 1 and np.array_equal(xs[-len(edge):], edge)
    assert planner.stats['gpu_launches'] > 0 and planner.stats['samples'] >= tree.size - 1


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_force_arrive_steer_matches_reference(cuda_device, dtype):
    """lq_steer_force = `_steer(ID, xtar, force_arrive=True)` (planner.py:294-303, 399-410) against the reference's
    own results on its boat_novice tree (tests/golden/make_golden_r2.py): targets it arrives at (np.allclose exit,
    ~2000 steps) and targets behind an obstacle (infeasible exit with the FPR cut)."""
    from gpu_util import make_engine
    g = load_golden('r2_extra'); prob = problem('demo_boat_novice')
    eng = make_engine(prob, dtype, node_capacity=len(g['fa_state']) + 8, wave_capacity=64)
    eng.load_nodes(g['fa_state'])
    off = 0
    for ID, xt, L in zip(g['fa_ids'], g['fa_targets'], g['fa_len']):
        L = int(L)
        xs, us = eng.steer_force(int(ID), xt, 20000)
        want_x = g['fa_x'][off:off + L]; want_u = g['fa_u'][off:off + L]
        off += L
        if dtype == 'f64':
            assert len(xs) == L
            assert np.abs(xs - want_x).max() < 1e-9 and np.abs(us - want_u).max() < 1e-6
        else:
            # float32: the arrival test |x - xtar| <= 1e-4 (1 + |xtar|) sits at float32 resolution of a 40 m
            # coordinate, so the exit step may move; the trajectory itself must agree while both are running
            if L < 100:
                assert len(xs) == L
            m = min(len(xs), L)
            assert m >= min(L, 20)
            assert np.abs(xs[:m] - want_x[:m]).max() < 2e-2 and np.abs(xs[:min(m, 50)] - want_x[:min(m, 50)]).max() < 1e-3
    eng.close()


def test_timeout_gives_closest_to_guide(cuda_device):
    """planner.py:311-328: no plan by max_time => the node nearest the guide (inf-buffer dims masked)."""
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f64', wave=16, seed=1, min_time=3, max_time=3, sys_time=WaveClock())
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=[0] * 6)
    assert ok and not planner.plan_reached_goal
    st = planner.tree.state
    d = prob.goal[None, :2] - st[:, :2]
    assert planner.node_seq[-1] == int(np.argmin((d * d).sum(1)))
    assert planner.T == len(planner.x_seq) * planner.dt
    assert np.allclose(planner.get_state(0.0), prob.x0)


def test_max_nodes_returns_false(cuda_device):
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f32', wave=256, max_nodes=300, min_time=1e6, max_time=1e6)
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=[0] * 6)
    assert ok is False and planner.tree.size == 301 and hasattr(planner, 'node_seq')


def test_kill_from_another_thread(cuda_device):
    import threading
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f32', wave=64, min_time=30, max_time=30, max_nodes=1e6)
    threading.Timer(0.3, planner.kill_update).start()
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    assert ok is False and planner.killed is False and planner.stats['elapsed'] < 10


def test_custom_xrand_gen_callable(cuda_device):
    prob = problem('demo_boat_novice')
    calls = []

    def gen(p):
        calls.append(p.tree.size)
        return np.array([5.0 + len(calls) % 7, 3.0, 0.0, 0.8, 0.0, 0.0])
    planner = prob.make_planner(dtype='f64', wave=4, min_time=5, max_time=5, sys_time=WaveClock())
    planner.update_plan(prob.x0, None, xrand_gen=gen)
    assert len(calls) == 4 * planner.stats['waves'] and planner.tree.size > 1
    with pytest.raises(ValueError):
        planner.update_plan(prob.x0, prob.sample_space, xrand_gen='nope')


def test_no_goal_and_validation(cuda_device):
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(goal0=None)
    assert planner.update_plan(prob.x0, prob.sample_space) is False
    assert np.array_equal(planner.get_state(1.0), prob.x0)
    planner.set_goal(prob.goal)
    with pytest.raises(ValueError):
        planner.update_plan(prob.x0, prob.sample_space[:3])
    with pytest.raises(ValueError):
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=[0.5, 0.5])


def test_riccati_provider_plans_pendulum(cuda_device):
    """C1 through the public API with the FD-Jacobian + Riccati provider."""
    prob = problem('c1')
    planner = prob.make_planner(dtype='f64', seed=4, min_time=0, max_time=30, max_nodes=60000)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    assert planner.tree.size > 50
    xs = np.array(planner.x_seq)
    assert np.all(np.abs(np.diff(xs[:, 0])) < 0.2)


def test_adaptive_horizon_wave1_reproduces_reference(cuda_device):
    """The adaptive-horizon heuristic (planner.py:418-425) through the engine, two consecutive plans with
    the carried-over horizon_iters, against the unmodified reference's trees."""
    g = load_golden('adaptive_boat_novice')
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f64', wave=1, rng='numpy', horizon=(0.1, 3), min_time=300, max_time=300,
                                sys_time=WaveClock())
    assert planner.hfactor == 2 and planner.horizon_iters == 1
    np.random.seed(0)
    np.random.rand(setup_draws('demo_boat_novice'))
    for k in range(2):
        planner.sys_time.t = -1
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
        tree = planner.tree
        assert tree.size == len(g['state%d' % k])
        assert tree.pID == g['pID%d' % k].tolist()
        assert [len(e) for e in tree.x_seq] == g['edge_len%d' % k].tolist()
        assert np.abs(tree.state - g['state%d' % k]).max() <= 1e-9
        assert planner.horizon_iters == int(g['horizon_iters%d' % k])
        assert planner.T == pytest.approx(float(g['T%d' % k]), abs=1e-12)


def test_adaptive_horizon_waves_match_oracle(cuda_device):
    """W = 48: every rollout starts from the wave-start horizon, votes applied in sample order at commit."""
    from gpu_util import make_engine, tree_arrays
    from lqrrt_b200._native import Engine
    from oracle.wave import WaveOracle
    prob = problem('demo_boat_novice')
    eng = Engine(prob.system, 30, 0.1, 0.5, prob.error_tol, node_capacity=4096, wave_capacity=48, dtype='f64',
                 device=0, hfactor=2, hspan=(1, 30))
    eng.set_goal(prob.goal, prob.goal_buffer)
    eng.set_sampling(prob.sample_space, prob.goal_bias, 10)
    eng.set_horizon(1)
    eng.reset(prob.x0)
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, (0.1, 3), 0.1, 0.5, prob.error_tol)
    orc.reset(prob.x0)
    hist = []
    for w in range(8):
        eng.wave_expand(3, w, 0, 48, None, True); eng.wave_commit(None, 1)
        st = eng.wave_stats()
        X, _ = orc.sample_wave(3, w, 0, 48, prob.sample_space, prob.goal_bias, 10)
        orc.expand_wave(X)
        assert (st.tree_size, st.horizon_iters) == (orc.size, orc.horizon_iters), w
        hist.append(st.horizon_iters)
    t = tree_arrays(eng)
    assert t['parent'].tolist() == orc.parent and np.abs(t['state'] - orc.state).max() <= 1e-9
    assert t['edge_len'][1:].tolist() == [len(e) for e in orc.x_seq[1:]]
    assert max(hist) == 30 and max(t['edge_len']) > 20


def test_replanning_pattern(cuda_device):
    """The ROS node's tree-chaining use (lqrrt_node.py:389-500): the same planner is asked for a new tree
    every cycle with a ~fixed time budget (specific_time), seeded from a state along the previous plan, and
    can be killed from another thread."""
    import time
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f32', wave_max=512, min_time=5, max_time=5, max_nodes=400000)
    x0 = prob.x0.copy()
    sizes = []
    for cycle in range(4):
        t0 = time.time()
        ok = planner.update_plan(x0, prob.sample_space, goal_bias=prob.goal_bias, specific_time=0.1)
        dt = time.time() - t0
        assert ok and 0.1 <= dt < 1.5, (ok, dt, planner.tree.size)
        assert np.allclose(planner.get_state(0.0), x0, atol=1e-5)
        sizes.append(planner.tree.size)
        x0 = np.array(planner.get_state(min(1.0, planner.T)), dtype=np.float64)     # next seed lies on the current plan
    assert min(sizes) > 100 and planner.tree.size <= planner.max_nodes + 1
