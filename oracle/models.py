"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package
(`lqrrt_b200/`); only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline legs may use it.

NumPy (float64) twins of the user callbacks on the reference's hot path,
restated from the reference demos and parameterised by the same spec records
the device models take (`lqrrt_b200.models`), so that one parameter set drives
both the checker and the thing checked.  Each twin cites the reference lines
it follows; `tests/test_oracle_vs_reference.py` pins every one of them
bit-for-bit against the reference's own functions imported from
/root/reference (when present) and against `tests/golden/`.

`make_callbacks(system, dt)` returns (dynamics, lqr, erf, is_feasible) with the
reference's signatures (planner.py:35-59).
"""
import numpy as np
import numpy.linalg as npl

from lqrrt_b200 import models as M


# ----------------------------------------------------------------------------
# dynamics  (a5)
# ----------------------------------------------------------------------------

def _wrap(a_from, a_to):
    """Signed angle from a_from to a_to on the circle, the reference's
    atan2(sg*c - cg*s, cg*c + sg*s) form (demo_boat_novice.py:112-117)."""
    c = np.cos(a_from); s = np.sin(a_from)
    cg = np.cos(a_to); sg = np.sin(a_to)
    return np.arctan2(sg * c - cg * s, cg * c + sg * s)


def boat_dynamics(p):
    """demo_boat_novice.py:45-75 / demo_boat_intermediate.py:48-100 /
    demo_boat_advanced.py:78-130, selected by the spec's flags."""
    invM, D_pos, D_neg = p.invM, p.D_pos, p.D_neg
    u_max, B, invB, thrust_max = p.u_max, p.B, p.invB, p.thrust_max

    def dynamics(x, u, dt):
        u = np.array(u, dtype=np.float64)
        c = np.cos(x[2]); s = np.sin(x[2])
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        D = np.where(x[3:] >= 0, D_pos, D_neg)
        if p.rudder != 0.0:                       # intermediate:71-77, advanced:101-108
            vw = R[:2, :2].dot(x[3:5])
            ang = np.arctan2(vw[1], vw[0])
            u[2] = u[2] + p.rudder * _wrap(x[2], ang)
        if p.clip == 'wrench':                    # novice:67-69
            for i in range(3):
                if abs(u[i]) > u_max[i]:
                    u[i] = u_max[i] * np.sign(u[i])
        else:                                     # advanced:111
            u = B.dot(np.clip(invB.dot(u), -thrust_max, thrust_max))
        xdot = np.concatenate((R.dot(x[3:]), invM * (u - D * x[3:])))
        xnext = x + xdot * dt
        if p.carlike:                             # intermediate:88-98, advanced:120-128
            if x[3] > 0:
                xnext[5] = np.clip(np.abs(xnext[3] / p.velmax_pos0), 0, 1) * xnext[5]
            elif x[3] < 0:
                xnext[5] = np.clip(np.abs(xnext[3] / p.velmax_neg0), 0, 1) * xnext[5]
            if xnext[3] < 0:
                xnext[3] = 0
        return xnext
    return dynamics


def car_dynamics(p):
    """demo_car.py:46-72."""
    def dynamics(x, u, dt):
        vwx = np.cos(x[2]) * x[3]
        vwy = np.sin(x[2]) * x[3]
        u = np.clip(u, p.u_lo, p.u_hi)
        xdot = np.concatenate(([vwx, vwy, x[4]], p.invM * (u - p.D * x[3:])))
        xnext = x + xdot * dt
        if xnext[3] < 0:
            xnext[3] = 0
        xnext[4] = np.clip(np.abs(xnext[3] / p.velmax0), 0, 1) * xnext[4]
        return xnext
    return dynamics


def pendulum2_dynamics(p):
    """Config C1: explicit Euler of
    omega' = (tau - b omega - m g l sin(theta)) / (m l^2)."""
    ml2 = p.mass * p.length * p.length
    mgl = p.mass * p.gravity * p.length

    def dynamics(x, u, dt):
        tau = np.clip(np.atleast_1d(u)[0], -p.tau_max, p.tau_max)
        wdot = (tau - p.damping * x[1] - mgl * np.sin(x[0])) / ml2
        return np.array([x[0] + x[1] * dt, x[1] + wdot * dt])
    return dynamics


def double_integrator_dynamics(p):
    """Config C5: p' = p + v dt, v' = v + u dt."""
    D = p.axes

    def dynamics(x, u, dt):
        a = np.clip(np.asarray(u, dtype=np.float64), -p.a_max, p.a_max)
        return np.concatenate((x[:D] + x[D:] * dt, x[D:] + a * dt))
    return dynamics


def make_dynamics(spec):
    return {M.SYS_BOAT: boat_dynamics, M.SYS_CAR: car_dynamics,
            M.SYS_PENDULUM2: pendulum2_dynamics,
            M.SYS_DOUBLE_INTEGRATOR: double_integrator_dynamics}[spec.kind](spec)


# ----------------------------------------------------------------------------
# lqr providers  (a6)
# ----------------------------------------------------------------------------

def pd_rotated_lqr(spec, n, m):
    """demo_boat_novice.py:86-103 (m=3) and demo_car.py:95-113 (m=2)."""
    kp, kd = spec.kp, spec.kd
    S = np.eye(n) if spec.S is None else np.array(spec.S, dtype=np.float64)
    if m == 3:
        def lqr(x, u):
            c = np.cos(x[2]); s = np.sin(x[2])
            R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
            return (S, np.hstack((kp.dot(R.T), kd)))
    elif m == 2:
        def lqr(x, u):
            w2b = np.array([[np.cos(x[2]), np.sin(x[2]), 0], [0, 0, 1]])
            return (S, np.hstack((kp.dot(w2b), kd)))
    else:
        raise ValueError("PD-rotated gain is defined for the boat (m=3) and car (m=2)")
    return lqr


def constant_lqr(spec, n, m):
    """demo_pendulum.py:119-126."""
    S = np.array(spec.S, dtype=np.float64); K = np.array(spec.K, dtype=np.float64)
    return lambda x, u: (S, K)


def fd_jacobians(dynamics, x, u, dt, eps):
    """Central-difference A = df/dx (n x n), B = df/du (n x m) of the
    discrete-time map x+ = dynamics(x, u, dt).  Not in the reference
    (SURVEY.md M2); the device twin is csrc/models.cuh fd_riccati."""
    n = len(x); m = len(u)
    A = np.zeros((n, n)); B = np.zeros((n, m))
    for j in range(n):
        xp = np.array(x, dtype=np.float64); xm = np.array(x, dtype=np.float64)
        xp[j] += eps; xm[j] -= eps
        A[:, j] = (dynamics(xp, np.array(u, dtype=np.float64), dt)
                   - dynamics(xm, np.array(u, dtype=np.float64), dt)) / (2 * eps)
    for j in range(m):
        up = np.array(u, dtype=np.float64); um = np.array(u, dtype=np.float64)
        up[j] += eps; um[j] -= eps
        B[:, j] = (dynamics(np.array(x, dtype=np.float64), up, dt)
                   - dynamics(np.array(x, dtype=np.float64), um, dt)) / (2 * eps)
    return A, B


def riccati_sweep(A, B, Q, R, steps):
    """`steps` backward iterations of the finite-horizon discrete Riccati
    recursion from P = Q:  G = R + B'PB,  K = G^-1 B'PA,  P = Q + A'P(A - BK).
    Returns (P_0, K_0); sign convention u = K (target - state)
    (planner.py:386-387)."""
    P = np.array(Q, dtype=np.float64)
    K = np.zeros((B.shape[1], A.shape[0]))
    for _ in range(steps):
        BtP = B.T.dot(P)
        K = npl.solve(R + BtP.dot(B), BtP.dot(A))
        P = Q + A.T.dot(P).dot(A - B.dot(K))
    return P, K


def fd_riccati_lqr(spec, dynamics, dt):
    def lqr(x, u):
        u = np.zeros(spec.R.shape[0]) + np.asarray(u, dtype=np.float64)
        A, B = fd_jacobians(dynamics, np.asarray(x, dtype=np.float64), u, dt, spec.eps)
        snap = float(getattr(spec, 'snap', 0.0))
        if snap > 0:                              # Jacobian entries on a grid (csrc/models.cuh FdRiccatiLqr::snapped)
            A = np.rint(A / snap) * snap; B = np.rint(B / snap) * snap
        return riccati_sweep(A, B, spec.Q, spec.R, spec.steps)
    return lqr


def make_lqr(spec, dyn_spec, dt):
    n, m = dyn_spec.nstates, dyn_spec.ncontrols
    if spec.kind == M.LQR_PD_ROTATED:
        return pd_rotated_lqr(spec, n, m)
    if spec.kind == M.LQR_CONSTANT:
        return constant_lqr(spec, n, m)
    if spec.kind == M.LQR_FD_RICCATI:
        return fd_riccati_lqr(spec, make_dynamics(dyn_spec), dt)
    raise ValueError("unknown lqr provider")


# ----------------------------------------------------------------------------
# erf  (a7)
# ----------------------------------------------------------------------------

def make_erf(spec):
    """demo_boat_novice.py:106-118; np.subtract when no angle dims
    (planner.py:87, so that erf_v takes the broadcast path of planner.py:585)."""
    dims = tuple(spec.angle_dims)
    if not dims:
        return np.subtract

    def erf(xgoal, x):
        e = xgoal - x
        for i in dims:
            e[i] = _wrap(x[i], xgoal[i])
        return e
    return erf


# ----------------------------------------------------------------------------
# feasibility  (a8)
# ----------------------------------------------------------------------------

def _cloud_world(p, x):
    dx, dy, dh = p.dims[0], p.dims[1], p.dims[2]
    c = np.cos(x[dh]); s = np.sin(x[dh])
    R = np.array([[c, -s], [s, c]])
    pos = np.array([x[dx], x[dy]])
    if p.flags & 1:      # demo_car.py:175 -- the stacked centre row also gets + x[:2]
        return pos + np.vstack((R.dot(p.verts).T, pos))
    return pos + R.dot(p.verts).T


def _prim_ok(p, x, u):
    k = p.kind
    if k == M.FEAS_STATE_BOX:          # demo_boat_advanced.py:211-213
        for i in range(len(p.lo)):
            if x[i] > p.hi[i] or x[i] < p.lo[i]:
                return False
        return True
    if k == M.FEAS_EFFORT_BOX:         # demo_pendulum.py:154-157
        return not np.any(np.abs(np.atleast_1d(u)) > p.hi)
    if k == M.FEAS_STATE_REGION:
        return not (np.all(x >= p.lo) and np.all(x <= p.hi))
    if k == M.FEAS_CIRCLES_POINT:      # demo_boat_novice.py:160-164
        pos = np.array([x[p.dims[0]], x[p.dims[1]]])
        for ob in p.obs:
            if npl.norm(pos - ob[:2]) <= p.inflate + ob[2]:
                return False
        return True
    if k == M.FEAS_CIRCLES_CLOUD:      # demo_boat_intermediate.py:198-210
        verts = _cloud_world(p, x)
        for ob in p.obs:
            if np.any(npl.norm(verts - ob[:2], axis=1) <= ob[2]):
                return False
        return True
    if k == M.FEAS_BOXES_CLOUD:
        verts = _cloud_world(p, x)
        for ob in p.obs:
            if np.any((np.abs(verts[:, 0] - ob[0]) <= ob[2]) & (np.abs(verts[:, 1] - ob[1]) <= ob[3])):
                return False
        return True
    if k == M.FEAS_OGRID_CLOUD:        # lqrrt_node.py:730-745
        dx, dy, dh = p.dims[0], p.dims[1], p.dims[2]
        c, s = np.cos(x[dh]), np.sin(x[dh])
        R = np.array([[c, -s], [s, c]])
        points = np.array([x[dx], x[dy]]) + R.dot(p.verts).T
        ind = (p.cpm * (points - p.origin)).astype(np.int64)
        try:
            vals = p.obs[ind[:, 1], ind[:, 0]]
        except IndexError:
            return False
        return bool(np.all(vals < p.threshold))
    if k == M.FEAS_SPHERES_POINT:
        pos = np.array([x[p.dims[0]], x[p.dims[1]], x[p.dims[2]]])
        for ob in p.obs:
            if npl.norm(pos - ob[:3]) <= p.inflate + ob[3]:
                return False
        return True
    raise ValueError("unknown feasibility primitive")


def make_is_feasible(spec):
    prims = list(spec.prims)

    def is_feasible(x, u):
        for p in prims:
            if not _prim_ok(p, x, u):
                return False
        return True
    return is_feasible


def make_is_feasible_batch(spec):
    """Vectorised over rows of X (B x n), U (B x m): same predicates, used by
    the wave oracle where per-row Python calls would take minutes."""
    prims = list(spec.prims)

    def feas(X, U):
        X = np.atleast_2d(X); U = np.atleast_2d(U)
        ok = np.ones(len(X), dtype=bool)
        for p in prims:
            k = p.kind
            if k == M.FEAS_STATE_BOX:
                nb = len(p.lo)
                ok &= ~np.any((X[:, :nb] > p.hi) | (X[:, :nb] < p.lo), axis=1)
            elif k == M.FEAS_EFFORT_BOX:
                ok &= ~np.any(np.abs(U) > p.hi, axis=1)
            elif k == M.FEAS_STATE_REGION:
                ok &= ~(np.all(X >= p.lo, axis=1) & np.all(X <= p.hi, axis=1))
            elif k in (M.FEAS_CIRCLES_POINT, M.FEAS_SPHERES_POINT):
                nd = 2 if k == M.FEAS_CIRCLES_POINT else 3
                if len(p.obs):
                    pos = X[:, list(p.dims[:nd])]
                    d = np.sqrt(((pos[:, None, :] - p.obs[None, :, :nd]) ** 2).sum(-1))
                    ok &= ~np.any(d <= p.inflate + p.obs[None, :, nd], axis=1)
            elif k in (M.FEAS_CIRCLES_CLOUD, M.FEAS_BOXES_CLOUD):
                if len(p.obs):
                    dx, dy, dh = p.dims[:3]
                    c = np.cos(X[:, dh]); s = np.sin(X[:, dh])
                    vx = X[:, dx, None] + c[:, None] * p.verts[0] - s[:, None] * p.verts[1]
                    vy = X[:, dy, None] + s[:, None] * p.verts[0] + c[:, None] * p.verts[1]
                    if p.flags & 1:
                        vx = np.hstack((vx, 2 * X[:, dx, None])); vy = np.hstack((vy, 2 * X[:, dy, None]))
                    for ob in p.obs:
                        if k == M.FEAS_CIRCLES_CLOUD:
                            hit = np.any(np.sqrt((vx - ob[0]) ** 2 + (vy - ob[1]) ** 2) <= ob[2], axis=1)
                        else:
                            hit = np.any((np.abs(vx - ob[0]) <= ob[2]) & (np.abs(vy - ob[1]) <= ob[3]), axis=1)
                        ok &= ~hit
            elif k == M.FEAS_OGRID_CLOUD:
                ok &= np.array([_prim_ok(p, x, u) for x, u in zip(X, U)], dtype=bool)
            else:
                raise ValueError("unknown feasibility primitive")
        return ok
    return feas


def make_is_feasible_allobs(spec):
    """`make_is_feasible` with the per-obstacle Python loop of the cloud / point primitives folded into one NumPy
    expression over all obstacles.  Same elementwise arithmetic (sqrt of a two- or three-term sum of squares, the
    same comparisons), so the verdicts are identical -- tests/test_oracle_golden.py checks that -- only ~50x
    faster on the 500-obstacle field, which is what lets the wave oracle follow 1024-sample waves in a test."""
    prims = list(spec.prims)

    def is_feasible(x, u):
        for p in prims:
            k = p.kind
            if k == M.FEAS_CIRCLES_CLOUD and len(p.obs):
                verts = _cloud_world(p, x)
                d = np.sqrt(((verts[None, :, :] - p.obs[:, None, :2]) ** 2).sum(-1))
                if np.any(d <= p.obs[:, None, 2]):
                    return False
            elif k == M.FEAS_BOXES_CLOUD and len(p.obs):
                verts = _cloud_world(p, x)
                if np.any((np.abs(verts[None, :, 0] - p.obs[:, None, 0]) <= p.obs[:, None, 2]) &
                          (np.abs(verts[None, :, 1] - p.obs[:, None, 1]) <= p.obs[:, None, 3])):
                    return False
            elif k in (M.FEAS_CIRCLES_POINT, M.FEAS_SPHERES_POINT) and len(p.obs):
                nd = 2 if k == M.FEAS_CIRCLES_POINT else 3
                pos = np.array([x[d] for d in p.dims[:nd]])
                if np.any(np.sqrt(((pos[None, :] - p.obs[:, :nd]) ** 2).sum(-1)) <= p.inflate + p.obs[:, nd]):
                    return False
            elif not _prim_ok(p, x, u):
                return False
        return True
    return is_feasible


def make_callbacks(system, dt, all_obstacles_at_once=False):
    dynamics = make_dynamics(system.dynamics)
    lqr = make_lqr(system.lqr, system.dynamics, dt)
    erf = make_erf(system.erf)
    is_feasible = (make_is_feasible_allobs if all_obstacles_at_once else make_is_feasible)(system.is_feasible)
    return dynamics, lqr, erf, is_feasible
