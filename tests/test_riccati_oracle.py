"""FD-Jacobian + finite-horizon Riccati provider (not in the reference; parity unpinned):
the NumPy restatement is anchored on SciPy's DARE solver as the horizon grows, and on exact Jacobians of
the linear double integrator."""
import numpy as np
import scipy.linalg

from lqrrt_b200 import models
from oracle import models as om


def test_fd_jacobians_exact_for_double_integrator():
    spec = models.DoubleIntegratorDynamics(3)
    dyn = om.make_dynamics(spec)
    x = np.arange(6.0); u = np.array([0.3, -0.2, 0.1])
    A, B = om.fd_jacobians(dyn, x, u, 0.1, 1e-3)
    A0 = np.eye(6); A0[:3, 3:] = 0.1 * np.eye(3)
    B0 = np.zeros((6, 3)); B0[3:] = 0.1 * np.eye(3)
    assert np.allclose(A, A0, atol=1e-10) and np.allclose(B, B0, atol=1e-10)


def test_long_horizon_converges_to_dare():
    spec = models.DoubleIntegratorDynamics(2)
    dyn = om.make_dynamics(spec)
    Q = np.eye(4); R = np.eye(2)
    A, B = om.fd_jacobians(dyn, np.zeros(4), np.zeros(2), 0.1, 1e-3)
    P, K = om.riccati_sweep(A, B, Q, R, 2000)
    Pinf = scipy.linalg.solve_discrete_are(A, B, Q, R)
    Kinf = np.linalg.solve(R + B.T.dot(Pinf).dot(B), B.T.dot(Pinf).dot(A))
    assert np.allclose(P, Pinf, rtol=1e-8) and np.allclose(K, Kinf, rtol=1e-8)
    # closed loop x+ = (A - B K) x is stable: u = K (target - x) drives x to the target
    assert np.max(np.abs(np.linalg.eigvals(A - B.dot(K)))) < 1


def test_pendulum_gain_swings_toward_target():
    sysm = models.pendulum2()
    dyn, lqr, erf, feas = om.make_callbacks(sysm, 0.01)
    S, K = lqr(np.array([0.3, 0.0]), np.zeros(1))
    assert S.shape == (2, 2) and K.shape == (1, 2)
    assert np.all(np.linalg.eigvalsh((S + S.T) / 2) > 0)
    assert K[0, 0] > 0 and K[0, 1] > 0
