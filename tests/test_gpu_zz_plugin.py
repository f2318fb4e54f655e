"""
User-compiled dynamics on the device (lqrrt_b200/plugin.py): the engine instantiated around CUDA source
supplied at run time must behave exactly like the engine around a compiled-in model with the same
arithmetic -- same callbacks, same tree bit for bit -- and a model the library has never seen (unicycle)
must match its NumPy statement and plan through the public Planner.
"""
import numpy as np
import pytest

import plugin_models as pm
from gpu_util import make_engine

pytestmark = pytest.mark.gpu


def _grow(prob, dtype, seed=5):
    planner = prob.make_planner(dtype=dtype, seed=seed, wave=128)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    t = planner.tree
    return t.size, list(t.pID), np.array(t.state), planner.stats['waves']


@pytest.mark.parametrize('dtype,riccati', [('f64', True), ('f64', False), ('f32', False)])
def test_user_double_integrator_equals_builtin(cuda_device, dtype, riccati):
    builtin, user = pm.dint2_pair(riccati)
    rng = np.random.RandomState(1)
    X = rng.randn(200, 4) * 3; U = rng.randn(200, 2) * 2
    ea = make_engine(builtin, 'f64'); eb = make_engine(user, 'f64')
    assert np.array_equal(ea.eval_dynamics(X, U, 0.1), eb.eval_dynamics(X, U, 0.1))
    Sa, Ka = ea.eval_lqr(X[:20], U[:20]); Sb, Kb = eb.eval_lqr(X[:20], U[:20])
    assert np.array_equal(Sa, Sb) and np.array_equal(Ka, Kb)
    a = _grow(builtin, dtype); b = _grow(user, dtype)
    assert a[0] == b[0] > 500 and a[3] == b[3]
    assert a[1] == b[1]
    assert np.array_equal(a[2].view(np.int64), b[2].view(np.int64))


def test_unicycle_matches_numpy_and_plans(cuda_device):
    prob = pm.unicycle_problem()
    eng = make_engine(prob, 'f64')
    rng = np.random.RandomState(2)
    X = rng.randn(300, 5) * np.array([20, 20, 3, 1, 0.5]); U = rng.randn(300, 2) * 4
    want = np.array([pm.unicycle_numpy(x, u, 0.1) for x, u in zip(X, U)])
    got = eng.eval_dynamics(X, U, 0.1)
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) <= 1e-12
    # spec call form (reference signature) goes through the same device code
    assert np.allclose(prob.system.dynamics(X[0], U[0], 0.1), want[0], rtol=1e-12, atol=1e-12)
    planner = prob.make_planner(dtype='f64', seed=1, wave=256)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    tree = planner.tree
    assert tree.size > 1000
    # every stored edge is a rollout of the user's own step from its parent under u = K erf (planner.py:386-390)
    for node in (1, tree.size // 2, tree.size - 1):
        xs, us = np.array(tree.x_seq[node]), np.array(tree.u_seq[node])
        x = tree.state[tree.pID[node]]
        for k in range(len(xs)):
            x = pm.unicycle_numpy(x, us[k], 0.1)
            assert np.max(np.abs(x - xs[k])) <= 1e-9
        assert np.allclose(tree.state[node], xs[-1], atol=1e-12)
