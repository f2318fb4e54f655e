"""
Verified recognition of plain-Python model callbacks.

The reference's demos define `dynamics / lqr / erf / is_feasible` as closures over module globals
(`invM, D_pos, D_neg, u_max, B, invB, thrust_max, magic_rudder, velmax_*, kp, kd, obs, vps, boat_length`,
e.g. demos/demo_boat_advanced.py:25-225).  A kernel cannot call them, but for the vehicles the engine
has compiled models of, the parameters can be read from those globals.  This module proposes candidate
specs from the globals and ACCEPTS one only if the device model (float64, evaluated through the C ABI)
agrees with the user's own functions on a few hundred random probes -- the user's code is the judge, so
a wrong guess can never silently change a planning problem.  If nothing verifies, the caller raises the
usual TypeError: there is still no CPU fallback.  The check is statistical: a constraint that none of the
probes exercises (a forbidden sliver thinner than ~1/500 of the probed box) cannot be seen; the probes are
therefore concentrated in, around and between the obstacles the globals list.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import models as M

PROBES = 256             # dynamics / lqr / erf probes
FEAS_PROBES = 2048       # is_feasible probes: thin forbidden regions need many (the check is statistical)
TOL = 1e-9


class NotRecognised(TypeError):
    pass


def _get(g, name, shape=None):
    if name not in g:
        return None
    try:
        a = np.array(g[name], dtype=np.float64)
    except Exception:
        return None
    if shape is not None and a.shape != tuple(shape):
        return None
    return a


# --------------------------------------------------------------------------------------------
# candidates
# --------------------------------------------------------------------------------------------

def dynamics_candidates(fn, n, m):
    g = getattr(fn, '__globals__', {})
    out = []
    if (n, m) == (6, 3):
        invM, Dp, Dn = _get(g, 'invM', (3,)), _get(g, 'D_pos', (3,)), _get(g, 'D_neg', (3,))
        if invM is None or Dp is None or Dn is None:
            return out
        vpos = _get(g, 'velmax_pos'); vneg = _get(g, 'velmax_neg')
        vp0 = float(np.ravel(vpos)[0]) if vpos is not None else 1.0
        vn0 = float(np.ravel(vneg)[0]) if vneg is not None else 1.0
        rudders = [0.0]
        if 'magic_rudder' in g:
            rudders.insert(0, float(g['magic_rudder']))
        rudders += [r for r in (5000.0, 4000.0) if r not in rudders]     # demo_boat_intermediate.py:77 hard-codes 5000
        B, invB, tmax = _get(g, 'B', (3, 4)), _get(g, 'invB', (4, 3)), _get(g, 'thrust_max', (4,))
        umax = _get(g, 'u_max', (3,))
        for rudder, carlike in itertools.product(rudders, (True, False)):
            if B is not None and invB is not None and tmax is not None:
                out.append(M.BoatDynamics(invM, Dp, Dn, B=B, invB=invB, thrust_max=tmax, clip='thruster',
                                          rudder=rudder, carlike=carlike, velmax_pos0=vp0, velmax_neg0=vn0))
            if umax is not None:
                out.append(M.BoatDynamics(invM, Dp, Dn, u_max=umax, clip='wrench', rudder=rudder, carlike=carlike,
                                          velmax_pos0=vp0, velmax_neg0=vn0))
    elif (n, m) == (5, 2):
        invM, D, umax, vel = _get(g, 'invM', (2,)), _get(g, 'D', (2,)), _get(g, 'u_max', (2,)), _get(g, 'velmax')
        if invM is not None and D is not None and umax is not None and vel is not None:
            out.append(M.CarDynamics(invM, D, [-umax[0] / 10, -umax[1]], umax, float(np.ravel(vel)[0])))   # demo_car.py:57
    return out


def lqr_candidates(fn, n, m, probe_x):
    g = getattr(fn, '__globals__', {})
    kp, kd = _get(g, 'kp', (m, m)), _get(g, 'kd', (m, m))
    try:
        S = np.array(fn(probe_x, np.zeros(m))[0], dtype=np.float64)
    except Exception:
        return []
    out = []
    if kp is not None and kd is not None and S.shape == (n, n) and (n, m) in ((6, 3), (5, 2)):
        out.append(M.PDRotatedLqr(kp, kd, S=S))
    try:
        K = np.array(fn(probe_x, np.zeros(m))[1], dtype=np.float64)
        if K.shape == (m, n) and S.shape == (n, n):
            out.append(M.ConstantLqr(S, K))
    except Exception:
        pass
    return out


def erf_candidates(fn, n, rng):
    if fn is np.subtract or fn is None:
        return [M.Erf()]
    dims = set()
    for _ in range(16):
        a = rng.uniform(-1, 1, n); b = rng.uniform(-1, 1, n)
        a2 = a + 2 * np.pi * rng.randint(-2, 3, n)                 # shifting a wrapped dim by whole turns changes nothing
        e1 = np.array(fn(np.copy(a), np.copy(b)), dtype=np.float64)
        e2 = np.array(fn(np.copy(a2), np.copy(b)), dtype=np.float64)
        for i in range(n):
            if a2[i] != a[i] and abs(e1[i] - e2[i]) < 1e-6:
                dims.add(i)
    return [M.Erf(sorted(dims))]


def feasibility_candidates(fn, n, m):
    g = getattr(fn, '__globals__', {})
    out = []
    obs = g.get('obs', None)
    try:
        obs = np.array(obs, dtype=np.float64).reshape(-1, 3) if obs is not None and len(obs) else np.zeros((0, 3))
    except Exception:
        obs = None
    box = []
    vp, vn = _get(g, 'velmax_pos_plan'), _get(g, 'velmax_neg_plan')
    if vp is not None and vn is not None and len(vp) == n - 3 and len(vn) == n - 3:
        lo = np.concatenate(([-np.inf] * 3, vn)); hi = np.concatenate(([np.inf] * 3, vp))
        box = [M.StateBox(lo, hi)]
    vps = _get(g, 'vps')
    if obs is not None:
        if vps is not None and vps.ndim == 2 and vps.shape[0] == 2:
            for dc in ((True, False) if n == 5 else (False, True)):
                out.append(M.Feasibility(box + [M.CirclesCloud(obs, vps, double_center=dc)]))
                if box:
                    out.append(M.Feasibility([M.CirclesCloud(obs, vps, double_center=dc)]))
        for key in ('boat_length', 'car_length'):
            if key in g:
                out.append(M.Feasibility(box + [M.CirclesPoint(obs, inflate=float(g[key]) / 2)]))
        out.append(M.Feasibility(box + [M.CirclesPoint(obs, inflate=0.0)]))
    umax_plan = g.get('umax_plan', None)
    if umax_plan is not None and np.ndim(umax_plan) == 0:
        out.append(M.Feasibility([M.EffortBox([float(umax_plan)] * m)]))
    out.append(M.Feasibility(box))
    return out, obs


# --------------------------------------------------------------------------------------------
# probes and verification (the user's functions are the reference; the device model is evaluated in float64)
# --------------------------------------------------------------------------------------------

def _probe_states(rng, n, obs, count=PROBES):
    X = np.zeros((count, n))
    X[:, :2] = rng.uniform(-10, 70, (count, 2))
    if n > 2:
        X[:, 2] = rng.uniform(-7, 7, count)
    if n > 3:
        X[:, 3:] = rng.uniform(-1.5, 1.5, (count, n - 3))
        X[::2, 3:] *= 0.1                              # half the probes inside any sensible speed limit
        X[::2, 3] = np.abs(X[::2, 3])
    if obs is not None and len(obs):
        k = min(count // 2, 4 * len(obs))
        pick = obs[rng.randint(0, len(obs), k)]
        ang = rng.uniform(0, 2 * np.pi, k)
        dist = (np.abs(pick[:, 2]) + 3.0) * rng.uniform(0.0, 1.6, k)      # in, around and outside every obstacle
        X[:k, 0] = pick[:, 0] + dist * np.cos(ang); X[:k, 1] = pick[:, 1] + dist * np.sin(ang)
    return X


def _agree(a, b, tol=TOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))))


def recognise(dynamics, lqr, erf, is_feasible, n, m, dt=0.1, seed=0):
    """Returns a verified models.System for plain callables, or raises NotRecognised."""
    rng = np.random.RandomState(seed)
    feas_c, obs = feasibility_candidates(is_feasible, n, m)
    X = _probe_states(rng, n, obs)
    U = rng.uniform(-3000, 3000, (len(X), m))
    XG = _probe_states(rng, n, None)
    dyn_c = dynamics_candidates(dynamics, n, m)
    if not dyn_c:
        raise NotRecognised("no compiled vehicle model matches the globals of `dynamics`")
    lqr_c = lqr_candidates(lqr, n, m, X[0])
    erf_c = erf_candidates(erf, n, rng)
    if not lqr_c:
        raise NotRecognised("`lqr` is not a rotated-PD or constant-gain provider the engine knows")

    want_dyn = np.array([dynamics(np.copy(x), np.copy(u), dt) for x, u in zip(X, U)], dtype=np.float64)
    want_K = np.array([lqr(x, u)[1] for x, u in zip(X[:64], U[:64])], dtype=np.float64)
    want_S = np.array([lqr(x, u)[0] for x, u in zip(X[:64], U[:64])], dtype=np.float64)
    want_e = np.array([erf(np.copy(a), np.copy(b)) for a, b in zip(XG, X)], dtype=np.float64)
    XF = np.vstack((X, _probe_states(rng, n, obs, FEAS_PROBES - len(X))))
    UF = np.vstack((U, rng.uniform(-3000, 3000, (len(XF) - len(X), m))))
    want_f = np.array([bool(is_feasible(x, u)) for x, u in zip(XF, UF)])

    def first(cands, check):
        for c in cands:
            try:
                if check(c):
                    return c
            except Exception:
                continue
        return None

    def mk(d, l, e, f):
        return M.System(d, l, e, f, dt_hint=dt, name='recognised')

    d_ok = first(dyn_c, lambda d: _agree(mk(d, lqr_c[0], erf_c[0], feas_c[-1]).evaluator().engine.eval_dynamics(X, U, dt), want_dyn))
    if d_ok is None:
        raise NotRecognised("the device vehicle models do not reproduce `dynamics` on {} probes".format(len(X)))

    def lqr_check(l):
        S, K = mk(d_ok, l, erf_c[0], feas_c[-1]).evaluator().engine.eval_lqr(X[:64], U[:64])
        return _agree(K, want_K) and _agree(S, want_S)
    l_ok = first(lqr_c, lqr_check)
    if l_ok is None:
        raise NotRecognised("no device lqr provider reproduces `lqr`")
    e_ok = first(erf_c, lambda e: _agree(mk(d_ok, l_ok, e, feas_c[-1]).evaluator().engine.eval_erf(XG, X), want_e, 1e-8))
    if e_ok is None:
        raise NotRecognised("`erf` is not a difference with wrapped angle dims")
    f_ok = first(feas_c, lambda f: bool(np.array_equal(mk(d_ok, l_ok, e_ok, f).evaluator().engine.eval_feasible(XF, UF), want_f)))
    if f_ok is None:
        raise NotRecognised("no combination of device feasibility primitives reproduces `is_feasible`")
    if want_f.all() and obs is not None and len(obs):
        raise NotRecognised("probes never hit an obstacle; refusing to certify `is_feasible`")
    return mk(d_ok, l_ok, e_ok, f_ok)
