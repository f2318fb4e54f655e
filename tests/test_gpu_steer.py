"""
K2 parity (level 1): device _steer against the reference's own _steer outputs (tests/golden) and the
oracle on random (node, target) pairs.  float64: same length, states to 1e-9.  float32: states to 1e-3
absolute over a full rollout, efforts 1e-3 relative to the gain scale, and a small budget of branch
flips (a feasibility or convergence test landing on the other side of its threshold).
"""
import numpy as np
import pytest

from conftest import DEMOS, load_golden, problem
from gpu_util import make_engine, steer
from oracle.wave import WaveOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
@pytest.mark.parametrize('name', DEMOS)
def test_steer_vs_reference_vectors(cuda_device, name, dtype):
    g = load_golden(name); prob = problem(name)
    eng = make_engine(prob, dtype)
    eng.load_nodes(g['state'], g['pID'])
    # loaded nodes get K = lqr(x, 0); the demo gains do not depend on u, so this equals the stored K
    xs, us, ln = steer(eng, g['hot_argmin'], g['hot_q'])
    flips = int((ln != g['hot_steer_len']).sum())
    assert flips == 0 if dtype == 'f64' else flips <= 1
    same = ln == g['hot_steer_len']
    for j in np.nonzero(same)[0]:
        L = int(ln[j])
        if L == 0:
            continue
        xt = 1e-9 if dtype == 'f64' else 1e-3
        ut = 1e-7 if dtype == 'f64' else 0.5
        assert np.abs(xs[j, :L] - g['hot_steer_x'][j, :L]).max() <= xt
        assert np.abs(us[j, :L] - g['hot_steer_u'][j, :L]).max() <= ut
        assert np.all(xs[j, L:] == 0)


@pytest.mark.parametrize('name,dtype', [('demo_boat_advanced', 'f64'), ('demo_boat_intermediate', 'f32'),
                                        ('c2', 'f64'), ('c3', 'f32'), ('c1', 'f64')])
def test_steer_random_pairs(cuda_device, name, dtype):
    prob = problem(name)
    rng = np.random.RandomState(17)
    ss = np.array(prob.sample_space); n = prob.nstates
    # nodes: short random walks out of x0 so that they are feasible, moving states
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)
    orc.reset(prob.x0)
    for w in range(4):
        X = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(24, n)
        orc.expand_wave(X)
    assert orc.size > 8
    eng = make_engine(prob, dtype)
    eng.load_nodes(orc.state, orc.parent, orc.cost)
    W = 96
    ids = rng.randint(0, orc.size, W)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(W, n)
    # K of a loaded node is lqr(x, 0); make the oracle use the same
    orc.K = [np.array(orc.lqr(x, np.zeros(prob.ncontrols))[1], dtype=np.float64) for x in orc.state]
    xs, us, ln = steer(eng, ids, Q)
    flips = 0; kept = 0
    for j in range(W):
        ox, ou = orc.steer(int(ids[j]), Q[j])
        if len(ox) != ln[j]:
            flips += 1
            continue
        if len(ox):
            kept += 1
            tol = 1e-8 if dtype == 'f64' else 2e-3
            if prob.system.lqr.kind == 3:
                tol = max(tol, 1e-5)
            assert np.abs(xs[j, :ln[j]] - np.array(ox)).max() <= tol, (j, np.abs(xs[j, :ln[j]] - np.array(ox)).max())
    assert flips <= (0 if dtype == 'f64' else 2), flips
    assert kept >= 5
