"""
State-independent metric of a state-dependent provider (LQ_OPT_WHITEN_METRIC, csrc/nearest.cuh whiten_apply): the
FD-Jacobian + Riccati provider on linear dynamics (config C5) returns the same S = lqr(x, 0).S (planner.py:348) for every
sample up to finite-difference noise, so the engine factors the probed mean S = U^T U once and searches a whitened copy
of the tree with the diagonal scan.  In float64 that is the per-sample-metric search to 1e-11, i.e. the oracle's answer;
the option off keeps one Riccati sweep per sample and the dense scan.  Nonlinear dynamics (the pendulum) are left alone.
"""
import numpy as np
import pytest

from conftest import problem
from gpu_util import make_engine, nearest
from lqrrt_b200 import _native as nat
from oracle.wave import WaveOracle

pytestmark = pytest.mark.gpu

OPT_WHITEN = 12


def _c5(axes=3, count=300):
    from lqrrt_b200 import configs
    return configs.c5_double_integrator(0, count=count, axes=axes, riccati=True)


def _oracle(prob):
    return WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)


@pytest.mark.parametrize('dtype,axes', [('f64', 3), ('f32', 3), ('f64', 6)])
def test_whitened_search_is_the_per_sample_metric_search(cuda_device, dtype, axes):
    prob = _c5(axes)
    n = prob.nstates
    rng = np.random.RandomState(4)
    ss = np.array(prob.sample_space)
    nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(1500, n)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(96, n)
    eng = make_engine(prob, dtype, node_capacity=2048, wave_capacity=128)
    eng.load_nodes(nodes)
    assert eng.get_option(OPT_WHITEN) == 0                      # nothing probed yet: per-sample metric
    i0, c0 = nearest(eng, Q)
    eng.set_sampling(prob.sample_space, prob.goal_bias, 10)    # probes S over the box; the loaded tree is whitened lazily
    assert eng.get_option(OPT_WHITEN) == 1
    i1, c1 = nearest(eng, Q)
    orc = _oracle(prob); orc.load(nodes, [-1] * len(nodes), K=[None] * len(nodes))
    tol = 1e-9 if dtype == 'f64' else 3e-2                      # float32: the per-sample S carries the FD noise itself
    for j in range(len(Q)):
        c = orc.costs(Q[j]); best = int(np.argmin(c))
        assert i1[j] == best or c[i1[j]] <= c[best] * (1 + tol)
        assert abs(c1[j] - c[i1[j]]) <= tol * max(1.0, c[best])
    if dtype == 'f64':
        assert np.array_equal(i0, i1) and np.abs(c0 - c1).max() <= 1e-8 * np.abs(c0).max()
    eng.set_option(OPT_WHITEN, 0)                               # back to one sweep per sample
    assert eng.get_option(OPT_WHITEN) == 0
    i2, c2 = nearest(eng, Q)
    assert np.array_equal(i2, i0) and np.array_equal(c2, c0)


def test_nonlinear_dynamics_keep_the_per_sample_metric(cuda_device):
    prob = problem('c1')                                        # pendulum: S depends on the angle
    eng = make_engine(prob, 'f64')
    eng.set_sampling(prob.sample_space, prob.goal_bias, 10)
    assert eng.get_option(OPT_WHITEN) == 0


def _grow(prob, dtype, opts, schedule='device', lag=1, nodes=2500, wave=64):
    planner = prob.make_planner(dtype=dtype, seed=5, max_nodes=nodes, min_time=1e9, max_time=1e9, wave=wave,
                                schedule=schedule, lag=lag)
    planner.engine_options = dict(opts)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    t = planner.tree
    out = dict(size=t.size, state=np.array(t.state), pID=list(t.pID), samples=planner.stats['samples'],
               whiten=planner._engine.get_option(OPT_WHITEN))
    planner._engine.close()
    return out


def test_whitened_plans_float64(cuda_device):
    prob = _c5(3)
    on = _grow(prob, 'f64', {})
    off = _grow(prob, 'f64', {OPT_WHITEN: 0})
    assert on['whiten'] == 1 and off['whiten'] == 0
    assert on['size'] == off['size'] == 2501 and on['pID'] == off['pID']          # same nearest nodes, hence the same tree
    assert np.abs(on['state'] - off['state']).max() <= 1e-9
    # every driver and the rollout kernel's own rescan search the same whitened copy with the same arithmetic: same bits
    for kw, opts in ((dict(schedule='host'), {}), (dict(lag=2), {}), (dict(), {nat.LQ_OPT_FORCE_RESCAN: 1})):
        other = _grow(prob, 'f64', opts, **kw)
        if kw.get('lag') == 2:
            base = _grow(prob, 'f64', {nat.LQ_OPT_FORCE_RESCAN: 1}, lag=2)
            assert other['pID'] == base['pID'] and other['state'].tobytes() == base['state'].tobytes()
        else:
            assert other['pID'] == on['pID'] and other['state'].tobytes() == on['state'].tobytes()


def test_whitened_plans_float32(cuda_device):
    prob = _c5(6, count=500)
    on = _grow(prob, 'f32', {}, nodes=6000, wave=256)
    off = _grow(prob, 'f32', {OPT_WHITEN: 0}, nodes=6000, wave=256)
    assert on['whiten'] == 1 and off['whiten'] == 0 and on['size'] == off['size'] == 6001
    # float32: the two searches differ within the finite-difference noise of S, the trees are different but alike
    assert abs(on['samples'] - off['samples']) <= 0.1 * off['samples']
    forced = _grow(prob, 'f32', {nat.LQ_OPT_FORCE_RESCAN: 1}, nodes=6000, wave=256)
    assert forced['pID'] == on['pID'] and forced['state'].tobytes() == on['state'].tobytes()


# ---- Jacobian snapping and the gain kept along a rollout (csrc/models.cuh FdRiccatiLqr::snapped, solve_coop) ----------
def test_snapped_gains_match_the_oracle(cuda_device):
    """lqr(x, u) of the C5 provider (snap = 0.025) against the NumPy restatement with the same snapping, saturated
    and unsaturated efforts; float32 lands on the same snapped Jacobian, so its gains are the float64 gains to 1e-4."""
    import oracle.models as om
    prob = _c5(3)
    rng = np.random.RandomState(7)
    n, m = prob.nstates, prob.ncontrols
    ss = np.array(prob.sample_space)
    X = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(24, n)
    U = 3.0 * (rng.rand(24, m) - 0.5); U[:6] = 0.0
    dyn, lqr, _, _ = om.make_callbacks(prob.system, prob.dt)
    want = [lqr(x, u) for x, u in zip(X, U)]
    for dtype, tol in (('f64', 1e-9), ('f32', 2e-4)):
        eng = make_engine(prob, dtype)
        S, K = eng.eval_lqr(X, U)
        for k in range(len(X)):
            assert np.abs(S[k] - want[k][0]).max() <= tol * np.abs(want[k][0]).max()
            assert np.abs(K[k] - want[k][1]).max() <= tol * max(1.0, np.abs(want[k][1]).max())
    assert len({np.round(w[1], 9).tobytes() for w in want}) > 1          # the saturation patterns give different gains


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_kept_gain_is_the_recomputed_gain(cuda_device, dtype):
    """A warp-per-sample rollout keeps its gain while the (snapped) Jacobian does not change; the thread-per-sample
    form runs the whole sweep after every recorded step.  Same edges, bit for bit."""
    from gpu_util import steer
    prob = _c5(3)
    rng = np.random.RandomState(11)
    n = prob.nstates
    ss = np.array(prob.sample_space)
    nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(64, n)
    nodes[:, n // 2:] *= 0.3
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(48, n)
    ids = rng.randint(0, 64, 48)
    eng = make_engine(prob, dtype, node_capacity=128, wave_capacity=64)
    eng.load_nodes(nodes)
    out = []
    for lanes in (32, 1):
        eng.set_tuning(lanes, 0)
        out.append(steer(eng, ids, Q))
    (x32, u32, l32), (x1, u1, l1) = out
    assert np.array_equal(l32, l1) and l32.max() >= 5
    assert x32.tobytes() == x1.tobytes() and u32.tobytes() == u1.tobytes()
