"""
The device twins of dynamics / lqr / erf / is_feasible (csrc/models.cuh) against the reference's own
outputs (tests/golden, made from the unmodified demos) and against the oracle on wider random inputs.
Tolerances: float64 engine 1e-12 relative (FMA contraction and CUDA libm vs NumPy), float32 engine 2e-5.
"""
import numpy as np
import pytest

from conftest import DEMOS, load_golden, problem
from gpu_util import make_engine
from oracle import models as om

pytestmark = pytest.mark.gpu

TOL = {'f64': 1e-11, 'f32': 3e-5}


def _close(a, b, tol, scale=None):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = np.maximum(1.0, np.abs(b)) if scale is None else scale
    return np.max(np.abs(a - b) / s) <= tol


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
@pytest.mark.parametrize('name', DEMOS)
def test_callbacks_vs_reference_vectors(cuda_device, name, dtype):
    g = load_golden(name); prob = problem(name)
    eng = make_engine(prob, dtype)
    tol = TOL[dtype]
    X, U, XG = g['cb_x'], g['cb_u'], g['cb_xg']
    assert _close(eng.eval_dynamics(X, U, 0.1), g['cb_dyn'], tol)
    S, K = eng.eval_lqr(X, U)
    assert _close(K, g['cb_K'], tol)
    assert _close(S, g['cb_S'], tol)
    assert _close(eng.eval_erf(XG, X), g['cb_erf'], tol * 10)
    ok = eng.eval_feasible(X, U)
    flips = int((ok != g['cb_feas']).sum())
    assert flips == 0 if dtype == 'f64' else flips <= 1, flips


@pytest.mark.parametrize('name', ['demo_boat_advanced', 'demo_car', 'c1', 'c2', 'c3', 'c4', 'c5small'])
def test_callbacks_vs_oracle_random(cuda_device, name):
    from lqrrt_b200 import configs
    prob = configs.c5_double_integrator(0, count=400, axes=3) if name == 'c5small' else problem(name)
    eng = make_engine(prob, 'f64')
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    rng = np.random.RandomState(5)
    ss = np.array(prob.sample_space)
    n, m = prob.nstates, prob.ncontrols
    X = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(300, n)
    X += 0.2 * rng.randn(300, n)
    XG = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(300, n)
    scale = {'c1': 12.0, 'c5small': 3.0}.get(name, 2000.0)
    U = scale * (rng.rand(300, m) - 0.5)
    want = np.array([dyn(np.copy(x), np.copy(u), prob.dt) for x, u in zip(X, U)])
    assert _close(eng.eval_dynamics(X, U, prob.dt), want, 1e-11)
    S, K = eng.eval_lqr(X[:40], U[:40])
    for i in range(40):
        So, Ko = lqr(X[i], U[i])
        # the finite-difference provider divides O(1e-16) rounding differences by 2*eps
        lim = 1e-6 if prob.system.lqr.kind == 3 else 1e-11
        assert _close(K[i], Ko, lim) and _close(S[i], So, lim)
    want = np.array([erf(np.copy(a), np.copy(b)) for a, b in zip(XG, X)])
    assert _close(eng.eval_erf(XG, X), want, 1e-10)
    want = np.array([bool(feas(x, u)) for x, u in zip(X, U)])
    got = eng.eval_feasible(X, U)
    assert (got != want).sum() == 0
    assert 0 < want.sum() < len(want)


def test_spec_call_evaluates_on_gpu(cuda_device):
    """A spec passed where the reference takes a function is itself callable (device-evaluated)."""
    prob = problem('demo_boat_novice')
    s = prob.system
    dyn, lqr, erf, feas = om.make_callbacks(s, prob.dt)
    x = np.array([3.0, 4.0, 0.7, 0.5, -0.1, 0.05]); u = np.array([100.0, -50.0, 20.0])
    assert np.allclose(s.dynamics(x, u, 0.1), dyn(np.copy(x), np.copy(u), 0.1), rtol=0, atol=1e-12)
    S, K = s.lqr(x, u)
    assert np.allclose(K, lqr(x, u)[1], atol=1e-12) and np.allclose(S, np.eye(6))
    assert np.allclose(s.erf(x + 1, x), erf(x + 1, np.copy(x)), atol=1e-12)
    assert s.is_feasible(x, u) == feas(x, u)


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_feasibility_full_size_fields(cuda_device, dtype):
    """The grid broad phase at BASELINE.json's full obstacle counts: C4's 500 circles vs the 28-vertex
    hull and C5's 5000 spheres, 20k random states each, against the vectorised oracle predicates."""
    from lqrrt_b200 import configs
    from oracle.models import make_is_feasible_batch
    rng = np.random.RandomState(8)
    for prob in (configs.c4_boat_advanced(0), configs.c5_double_integrator(0, riccati=False)):
        eng = make_engine(prob, dtype, node_capacity=16, wave_capacity=16)
        ss = np.array(prob.sample_space)
        n, m = prob.nstates, prob.ncontrols
        X = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(20000, n)
        if n == 6:
            X[:, 2] = 7 * (rng.rand(20000) - 0.5)
            X[:, 0:2] = -10 + 220 * rng.rand(20000, 2)          # also outside the obstacle field's bounds
        U = np.zeros((20000, m))
        want = make_is_feasible_batch(prob.system.is_feasible)(X, U)
        got = eng.eval_feasible(X, U)
        flips = int((got != want).sum())
        assert flips <= (0 if dtype == 'f64' else 4), flips
        assert 0.05 < want.mean() < 0.95


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_occupancy_grid_primitive(cuda_device, dtype):
    """The ROS node's is_feasible (lqrrt_node.py:719-745): hull vertices truncated to cell indices, hit if any
    cell >= threshold, IndexError (index past the end) => infeasible, negative indices wrap as in NumPy."""
    from lqrrt_b200 import models, configs
    from oracle.models import make_is_feasible
    rng = np.random.RandomState(4)
    grid = (rng.rand(120, 160) < 0.002) * 100.0
    grid[40:60, 50:52] = 100.0
    base = configs.demo_boat_novice(0)
    vps = models.vertex_cloud(models.BOAT_LENGTH, models.BOAT_WIDTH, 0.15, 0.4)
    og = models.OccupancyGrid(grid, cpm=2.5, origin=[-5.0, -8.0], threshold=90, verts=vps)
    system = models.System(base.system.dynamics, base.system.lqr, models.Erf([2]), models.Feasibility([og]))
    prob = configs.Problem('ogrid', system, base.x0, base.goal, base.goal_buffer, base.error_tol, base.sample_space,
                           base.goal_bias, base.horizon, base.dt, base.FPR)
    eng = make_engine(prob, dtype)
    X = np.zeros((3000, 6))
    X[:, 0] = rng.uniform(-12, 66, 3000); X[:, 1] = rng.uniform(-14, 46, 3000); X[:, 2] = rng.uniform(-4, 4, 3000)
    U = np.zeros((3000, 3))
    feas = make_is_feasible(system.is_feasible)
    want = np.array([feas(x, u) for x, u in zip(X, U)])
    got = eng.eval_feasible(X, U)
    flips = int((got != want).sum())
    assert flips <= (0 if dtype == 'f64' else 3), flips
    assert 0.05 < want.mean() < 0.9
    outside = (X[:, 0] > 62) | (X[:, 1] > 43)                     # past the grid's end: IndexError => infeasible
    assert not got[outside].any()
