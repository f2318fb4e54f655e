"""
Direct pin of the oracle against the unmodified reference imported from /root/reference (build
container only; skipped on the GPU box, where tests/golden/ stands in).  More seeds and the pruning
flag than the golden files carry.
"""
import numpy as np
import pytest

import ref_shim
from conftest import problem, setup_draws
from oracle import models as om
from oracle.seqplanner import SeqPlanner, IterationClock

pytestmark = pytest.mark.skipif(not ref_shim.have_reference(), reason="/root/reference not present")

CASES = [('demo_boat_novice', 2, 0.5), ('demo_car', 5, 0), ('demo_boat_intermediate', 2, 0.5),
         ('demo_boat_advanced', 2, 0.9)]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('seed,pruning', [(1, True), (2, False)])
def test_tree_bit_identical(case, seed, pruning):
    name, horizon, FPR = case
    K = 250
    rp, ns = ref_shim.run_reference_demo(name, seed, K, dict(horizon=horizon, dt=0.1, FPR=FPR), pruning=pruning)
    prob = problem(name, seed)
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, prob.nstates, prob.ncontrols, prob.goal_buffer, feas, prob.horizon, prob.dt,
                    prob.FPR, prob.error_tol, erf, min_time=K, max_time=K, goal0=prob.goal,
                    sys_time=IterationClock())
    np.random.seed(seed)
    np.random.rand(setup_draws(name))
    sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10, pruning=pruning)
    assert sp.tree.size == rp.tree.size
    assert np.array_equal(sp.tree.state, rp.tree.state)
    assert sp.tree.pID == rp.tree.pID
    assert sp.T == rp.T
    assert sp.plan_reached_goal == rp.plan_reached_goal
    assert np.array_equal(np.array(sp.x_seq), np.array(rp.x_seq))


def test_reference_validation_messages_match():
    """The host mirror raises the reference's ValueErrors."""
    ref = ref_shim.import_reference()
    import lqrrt_b200 as mine
    from lqrrt_b200 import models
    for mod, feas in ((ref, lambda x, u: True), (mine, models.Feasibility())):
        with pytest.raises(ValueError) as e1:
            mod.Constraints(3, 1, [1, 2], feas)
        assert "goal_buffer must have same dimensionality" in str(e1.value)
        with pytest.raises(ValueError) as e2:
            mod.Constraints(2, 1, [1, 2], 5)
        assert "Expected is_feasible to be a function" in str(e2.value)
