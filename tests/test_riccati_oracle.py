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


def test_snapped_jacobians_are_the_exact_ones_on_the_double_integrator():
    """`FDRiccatiLqr(snap=dt / 4)` (config C5): every true Jacobian entry of the saturating double integrator lies on that
    grid -- 1, dt, 0 -- so snapping turns the differenced Jacobian into the analytic one exactly, at any state and for
    every saturation pattern, and the gains are those of the analytic Jacobian (this is what the float32 engine relies on
    to see the float64 Jacobian).  Where a central difference straddles the saturation (|u| within eps of a_max) the
    slope is anything between 0 and dt; it lands on one of the five grid values in between."""
    dt, axes = 0.1, 3
    system = models.double_integrator(axes, a_max=1.0, snap=dt / 4)
    dyn, lqr, _, _ = om.make_callbacks(system, dt)
    spec = system.lqr
    rng = np.random.RandomState(3)
    patterns = set()
    for trial in range(40):
        x = 200.0 * (rng.rand(2 * axes) - 0.5)                     # large |x|: where differencing is noisiest
        u = 4.0 * (rng.rand(axes) - 0.5)
        if trial == 0:
            u[:] = 0.0
        if trial == 1:
            u[0] = 1.0 + 0.4e-3                                    # the difference straddles the saturation: slope 0.3 dt
        A, B = om.fd_jacobians(dyn, x, u, dt, spec.eps)
        As = np.rint(A / spec.snap) * spec.snap; Bs = np.rint(B / spec.snap) * spec.snap
        A0 = np.eye(2 * axes); A0[:axes, axes:] = dt * np.eye(axes)
        straddle = (np.abs(u) + spec.eps > 1.0) & (np.abs(u) - spec.eps < 1.0)
        slope = np.where(np.abs(u) + spec.eps <= 1.0, 1.0, 0.0)
        for i in np.nonzero(straddle)[0]:                          # (a_max - (|u| - eps)) / (2 eps), snapped
            slope[i] = np.rint((1.0 - (abs(u[i]) - spec.eps)) / (2 * spec.eps) * dt / spec.snap) * spec.snap / dt
        B0 = np.zeros((2 * axes, axes)); B0[axes:] = dt * np.diag(slope)
        assert np.abs(As - A0).max() < 1e-15 and np.abs(Bs - B0).max() < 1e-15
        assert trial != 1 or abs(Bs[axes, 0] - 0.025) < 1e-15
        S, K = lqr(x, u)
        S0, K0 = om.riccati_sweep(A0, B0, spec.Q, spec.R, spec.steps)
        assert np.abs(S - S0).max() < 1e-12 and np.abs(K - K0).max() < 1e-12
        patterns.add(tuple(slope))
    assert len(patterns) > 4                                       # several saturation patterns, each with its own gain


def test_snap_off_is_the_plain_provider():
    sysm = models.pendulum2()
    assert sysm.lqr.snap == 0.0 and len(sysm.lqr.blob(2, 1)) == 2 * 2 + 1 + 3
    with np.testing.assert_raises(ValueError):
        models.FDRiccatiLqr(np.eye(2), np.eye(1), snap=-1.0)
