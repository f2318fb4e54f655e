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
    # the wave oracle's all-obstacles-at-once form of is_feasible gives the same verdicts, also right on obstacle rims
    fast = om.make_callbacks(prob.system, prob.dt, all_obstacles_at_once=True)[3]
    rng = np.random.RandomState(1)
    X = np.vstack((g['cb_x'], g['state'][rng.randint(0, len(g['state']), 400)] + 0.5 * rng.randn(400, g['state'].shape[1])))
    obs = g['obs'][g['obs'][:, 0] > -1000]
    if len(obs):
        rim = X[:200].copy()
        pick = obs[rng.randint(0, len(obs), 200)]
        ang = 2 * np.pi * rng.rand(200)
        rim[:, 0] = pick[:, 0] + (pick[:, 2] + 3 * rng.rand(200)) * np.cos(ang)
        rim[:, 1] = pick[:, 1] + (pick[:, 2] + 3 * rng.rand(200)) * np.sin(ang)
        X = np.vstack((X, rim))
    u0 = np.zeros(prob.ncontrols)
    verdicts = [bool(feas(x, u0)) for x in X]
    assert verdicts == [bool(fast(x, u0)) for x in X]
    assert 0 < sum(verdicts) < len(verdicts)


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


def test_force_arrive_steer_bit_exact():
    """`_steer(ID, xtar, force_arrive=True)` (planner.py:294-303, 399-410) of the reference on its own boat_novice
    tree (tests/golden/make_golden_r2.py): targets it reaches and targets behind an obstacle."""
    g = load_golden('r2_extra')
    sp, prob = _run_seq('demo_boat_novice', 400)
    assert np.array_equal(sp.tree.state, g['fa_state'])
    off = 0
    for ID, xt, L in zip(g['fa_ids'], g['fa_targets'], g['fa_len']):
        xs, us = sp.steer_force(int(ID), xt, max_steps=10 ** 6)
        assert len(xs) == int(L)
        assert np.array_equal(np.array(xs).reshape(-1, 6), g['fa_x'][off:off + L])
        assert np.array_equal(np.array(us).reshape(-1, 3), g['fa_u'][off:off + L])
        off += int(L)
    assert g['fa_len'].max() > 1000 and g['fa_len'].min() < 100      # both exits are exercised


def test_diagonal_non_identity_metric_bit_exact():
    """`_costs_to_go` with the ROS car's S = diag(1, 1, 1, 0, 0) (behaviors/car.py:65) on the demo_car tree."""
    g = load_golden('r2_extra')
    prob = problem('demo_car')
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    S = g['ds_S']
    erf_v = make_erf_v(erf)
    for j, q in enumerate(g['ds_q']):
        c = costs_to_go(lambda x, u: (S, lqr(x, u)[1]), erf_v, 2, g['ds_state'], q)
        assert int(np.argmin(c)) == int(g['ds_argmin'][j])
        assert c.min() == g['ds_mincost'][j]
