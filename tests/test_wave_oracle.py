"""The wave-batched oracle with W = 1 is the sequential oracle (which is pinned bit-for-bit to the
reference): same samples => same tree, same ignores, same best plan."""
import numpy as np
import pytest

from conftest import problem
from oracle import models as om
from oracle.seqplanner import SeqPlanner, IterationClock
from oracle.wave import WaveOracle
from lqrrt_b200 import waves


@pytest.mark.parametrize('name,pruning', [('demo_boat_novice', True), ('demo_car', False)])
def test_wave1_equals_sequential(name, pruning):
    """The car's saturated dynamics create exactly duplicated nodes, hence exact cost ties.  With pruning
    the reference picks among tied nodes by np.argsort's unstable (and machine-dependent) order
    (planner.py:240-245); the engine and the wave oracle always take the lowest index, as np.argmin does
    (planner.py:247).  So the car is compared on the argmin path, the duplicate-free boat on the pruned one."""
    prob = problem(name)
    K = 300
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, prob.nstates, prob.ncontrols, prob.goal_buffer, feas, prob.horizon, prob.dt,
                    prob.FPR, prob.error_tol, erf, min_time=K, max_time=K, goal0=prob.goal, sys_time=IterationClock())
    trace = []
    np.random.seed(4)
    sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10, trace=trace, pruning=pruning)
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)
    orc.reset(prob.x0)
    for xrand, near, L in trace:
        added, ids, edges = orc.expand_wave(xrand[None, :], use_ignore=pruning)
        assert int(ids[0]) == near and len(edges[0][0]) == L
    assert orc.size == sp.tree.size
    assert np.array_equal(orc.state, sp.tree.state)
    assert orc.parent == sp.tree.pID
    assert sorted(np.nonzero(orc.ignore)[0].tolist()) == sorted(int(i) for i in sp.ignores)
    assert sp.plan_reached_goal or name != 'demo_car'
    if sp.plan_reached_goal:          # otherwise sp.T describes the closest-to-guide plan (planner.py:322)
        assert orc.T == sp.T
        assert orc.climb(orc.best_node) == sp.node_seq
    else:
        assert np.isinf(orc.T)


def test_philox_sampler_follows_reference_rules():
    """planner.py:201-211 on the Philox stream: inside the box, goal-snapped dims move together, feasible
    unless the tries ran out."""
    prob = problem('demo_boat_novice')
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)
    X, used = orc.sample_wave(3, 0, 0, 4000, prob.sample_space, prob.goal_bias, 10)
    ss = np.array(prob.sample_space)
    assert np.all(X >= ss[:, 0] - 1e-12) and np.all(X <= ss[:, 1] + 1e-12)
    snapped = (X[:, 0] == prob.goal[0])
    assert np.array_equal(snapped, X[:, 1] == prob.goal[1])
    assert 0.4 < snapped.mean() < 0.6
    ok = orc.feas_batch(X, np.zeros((len(X), 3)))
    assert ok[used < 10].all() and used.min() == 1
    # any split of the wave draws the same samples (what makes sharding invisible)
    Xa, _ = orc.sample_wave(3, 0, 0, 1500, prob.sample_space, prob.goal_bias, 10)
    Xb, _ = orc.sample_wave(3, 0, 1500, 2500, prob.sample_space, prob.goal_bias, 10)
    assert np.array_equal(np.vstack((Xa, Xb)), X)


def test_partition_and_wave_size():
    for W in (1, 7, 64, 8191, 8192):
        for G in (1, 2, 3, 8):
            parts = [waves.partition(W, r, G) for r in range(G)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == W
            for (f0, c0), (f1, _) in zip(parts[:-1], parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) <= waves.local_capacity(W, G)
    assert waves.wave_size(1) == 32 and waves.wave_size(500) == 500 and waves.wave_size(10 ** 6) == 8192
    assert waves.wave_size(10 ** 6, fixed=1) == 1


def test_wave1_equals_sequential_with_adaptive_horizon():
    prob = problem('demo_boat_novice')
    K = 250
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, 6, 3, prob.goal_buffer, feas, (0.1, 3), 0.1, 0.5, prob.error_tol, erf,
                    min_time=K, max_time=K, goal0=prob.goal, sys_time=IterationClock())
    trace = []
    np.random.seed(9)
    sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10, trace=trace)
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, (0.1, 3), 0.1, 0.5, prob.error_tol)
    orc.reset(prob.x0)
    seen = set()
    for xrand, near, L in trace:
        added, ids, edges = orc.expand_wave(xrand[None, :])
        assert int(ids[0]) == near and len(edges[0][0]) == L
        seen.add(orc.horizon_iters)
    assert orc.horizon_iters == sp.horizon_iters and max(seen) == 30      # one rollout can double 1 -> 30
    assert np.array_equal(orc.state, sp.tree.state) and orc.parent == sp.tree.pID
