"""
Pins the oracle (oracle/models.py, oracle/seqplanner.py) to outputs of the unmodified reference that
tests/golden/make_golden.py recorded (the reference itself cannot travel to the GPU box).  Float64,
bit-exact unless stated.
"""
import numpy as np
import pytest

from conftest import DEMOS, load_golden, problem, setup_draws
from oracle import models as om
from oracle.seqplanner import SeqPlanner, IterationClock, costs_to_go, make_erf_v


@pytest.mark.parametrize('name', DEMOS)
def test_callbacks_bit_exact(name):
    g = load_golden(name); prob = problem(name)
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    for i in range(len(g['cb_x'])):
        x, u, xg = g['cb_x'][i], g['cb_u'][i], g['cb_xg'][i]
        assert np.array_equal(dyn(np.copy(x), np.copy(u), 0.1), g['cb_dyn'][i])
        S, K = lqr(x, u)
        assert np.array_equal(np.asarray(K, dtype=np.float64), g['cb_K'][i])
        assert np.array_equal(np.asarray(S, dtype=np.float64), g['cb_S'][i])
        assert np.array_equal(erf(np.copy(xg), np.copy(x)), g['cb_erf'][i])
        assert bool(feas(x, u)) == bool(g['cb_feas'][i])
    assert 0 < g['cb_feas'].sum() < len(g['cb_feas'])        # both verdicts are exercised


@pytest.mark.parametrize('name', DEMOS)
def test_obstacle_field_matches_demo(name):
    g = load_golden(name); prob = problem(name)
    assert np.array_equal(np.asarray(prob.obs, dtype=np.float64).reshape(g['obs'].shape), g['obs'])


def _run_seq(name, K):
    prob = problem(name)
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, prob.nstates, prob.ncontrols, prob.goal_buffer, feas, prob.horizon, prob.dt,
                    prob.FPR, prob.error_tol, erf, min_time=K, max_time=K, goal0=prob.goal,
                    sys_time=IterationClock())
    np.random.seed(0)
    np.random.rand(setup_draws(name))
    sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    return sp, prob


@pytest.mark.parametrize('name', DEMOS)
def test_sequential_plan_bit_exact(name):
    """Parity level 2 of the oracle itself: same seed, same fake clock, 400 iterations."""
    g = load_golden(name)
    sp, prob = _run_seq(name, int(g['K']))
    assert sp.tree.size == len(g['state'])
    assert np.array_equal(sp.tree.state, g['state'])
    assert sp.tree.pID == g['pID'].tolist()
    assert [len(e) for e in sp.tree.x_seq] == g['edge_len'].tolist()
    assert sp.T == float(g['T'])
    assert sp.plan_reached_goal == bool(g['reached'])
    assert sp.node_seq == g['node_seq'].tolist()
    assert np.array_equal(np.array(sp.x_seq), g['x_seq'])
    assert np.array_equal(np.array(sp.u_seq), g['u_seq'])


@pytest.mark.parametrize('name', DEMOS)
def test_hot_path_functions(name):
    """_costs_to_go / selection / _steer on the reference's final tree."""
    g = load_golden(name)
    sp, prob = _run_seq(name, int(g['K']))
    erf_v = make_erf_v(sp.erf)
    for j, q in enumerate(g['hot_q']):
        c = costs_to_go(sp.lqr, erf_v, sp.ncontrols, sp.tree.state, q)
        assert int(np.argmin(c)) == int(g['hot_argmin'][j])
        assert c.min() == g['hot_mincost'][j]
        xs, us = sp.steer(int(g['hot_argmin'][j]), q)
        L = int(g['hot_steer_len'][j])
        assert len(xs) == L
        if L:
            assert np.array_equal(np.array(xs), g['hot_steer_x'][j, :L])
            assert np.array_equal(np.array(us), g['hot_steer_u'][j, :L])
    assert g['hot_steer_len'].max() > 0


def test_adaptive_horizon_two_plans_bit_exact():
    """planner.py:418-425, 538-547: horizon=(0.1, 3) as the ROS behaviours use it; horizon_iters is carried
    from the first plan into the second."""
    g = load_golden('adaptive_boat_novice')
    prob = problem('demo_boat_novice')
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    clock = IterationClock()
    sp = SeqPlanner(dyn, lqr, 6, 3, prob.goal_buffer, feas, (0.1, 3), 0.1, 0.5, prob.error_tol, erf,
                    min_time=300, max_time=300, goal0=prob.goal, sys_time=clock)
    np.random.seed(0)
    np.random.rand(setup_draws('demo_boat_novice'))
    for k in range(2):
        clock.t = -1
        sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
        assert np.array_equal(sp.tree.state, g['state%d' % k])
        assert sp.tree.pID == g['pID%d' % k].tolist()
        assert [len(e) for e in sp.tree.x_seq] == g['edge_len%d' % k].tolist()
        assert sp.horizon_iters == int(g['horizon_iters%d' % k])
        assert sp.T == float(g['T%d' % k])
    assert g['edge_len0'].max() > 20          # edges longer than the fixed-horizon demos' 20 steps occur
