"""
Device model specifications: the parameter side of the user callbacks
``dynamics / lqr / erf / is_feasible`` of the reference planner contract
(reference: lqrrt/planner.py:35-59, lqrrt/constraints.py:31-61).

The reference takes arbitrary Python closures; a CUDA kernel cannot call
Python, so this package takes *specs*: plain parameter records that select one
of the compiled device models in ``csrc/models.cuh`` and carry its constants.
A spec holds parameters only -- there is no CPU implementation of the model
arithmetic anywhere in this package.  Every spec is nevertheless *callable*
with the reference's callback signature (so it can be passed positionally to
``Planner`` / ``Constraints`` and used by demo code such as the post-plan
tracking simulation); the call is evaluated on the GPU through the C-ABI
(``lq_eval_*`` in include/lqrrt_b200.h) and fails loudly without it.

Blob layouts below are the wire format consumed by ``csrc/engine.cu``
(`unpack_*`); keep the two in sync.
"""
from __future__ import annotations

import numpy as np

# ---- kind ids (mirrors include/lqrrt_b200.h) --------------------------------
SYS_BOAT = 1
SYS_CAR = 2
SYS_PENDULUM2 = 3
SYS_DOUBLE_INTEGRATOR = 4

LQR_PD_ROTATED = 1
LQR_CONSTANT = 2
LQR_FD_RICCATI = 3

FEAS_STATE_BOX = 1
FEAS_EFFORT_BOX = 2
FEAS_STATE_REGION = 3
FEAS_CIRCLES_POINT = 4
FEAS_CIRCLES_CLOUD = 5
FEAS_BOXES_CLOUD = 6
FEAS_SPHERES_POINT = 7
FEAS_OGRID_CLOUD = 8

MAX_N = 12
MAX_M = 6


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.array(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape {}, got {}".format(tuple(shape), a.shape))
    return a


class _Spec(object):
    """Base of all specs: remembers the system it was bundled with so that a
    bare call (``spec(x, u)``) can build a one-off device evaluator."""
    _bundle = None

    def _evaluator(self):
        if self._bundle is None:
            raise RuntimeError(
                "{} is not attached to a lqrrt_b200.models.System; build the bundle "
                "first (models.System(dynamics=..., lqr=..., erf=..., is_feasible=...))"
                .format(type(self).__name__))
        return self._bundle.evaluator()


# =============================================================================
# dynamics
# =============================================================================

class Dynamics(_Spec):
    kind = 0
    nstates = 0
    ncontrols = 0

    def blob(self):
        raise NotImplementedError

    def __call__(self, x, u, dt):
        """xnext = dynamics(x, u, dt), evaluated on the device in float64."""
        return self._evaluator().dynamics(x, u, dt)


class BoatDynamics(Dynamics):
    """Planar boat, state [x, y, h, vx, vy, vh], wrench [ux, uy, uh]
    (reference arithmetic: demos/demo_boat_novice.py:45-75,
    demo_boat_intermediate.py:48-100, demo_boat_advanced.py:78-130).

    clip='wrench'   per-axis |u_i| <= u_max_i              (novice / intermediate)
    clip='thruster' u = B clip(invB u, +-thrust_max)       (advanced)
    rudder          gain of the heading "magic rudder" torque (0 disables)
    carlike         no-turn-in-place scaling + no-reverse clamp
    """
    kind = SYS_BOAT
    nstates = 6
    ncontrols = 3

    def __init__(self, invM, D_pos, D_neg, u_max=None, B=None, invB=None, thrust_max=None,
                 clip='wrench', rudder=0.0, carlike=False, velmax_pos0=1.0, velmax_neg0=1.0):
        self.invM = _f64(invM, (3,))
        self.D_pos = _f64(D_pos, (3,))
        self.D_neg = _f64(D_neg, (3,))
        if clip not in ('wrench', 'thruster'):
            raise ValueError("clip must be 'wrench' or 'thruster'")
        self.clip = clip
        if clip == 'wrench':
            self.u_max = _f64(u_max, (3,))
            self.B = np.zeros((3, 4)); self.invB = np.zeros((4, 3)); self.thrust_max = np.zeros(4)
        else:
            self.u_max = np.zeros(3)
            self.B = _f64(B, (3, 4)); self.invB = _f64(invB, (4, 3))
            self.thrust_max = _f64(thrust_max, (4,))
        self.rudder = float(rudder)
        self.carlike = bool(carlike)
        self.velmax_pos0 = float(velmax_pos0)
        self.velmax_neg0 = float(velmax_neg0)

    def blob(self):
        return np.concatenate((self.invM, self.D_pos, self.D_neg, self.u_max,
                               self.B.ravel(), self.invB.ravel(), self.thrust_max,
                               [self.rudder, self.velmax_pos0, self.velmax_neg0,
                                1.0 if self.clip == 'thruster' else 0.0,
                                1.0 if self.carlike else 0.0,
                                1.0 if self.rudder != 0.0 else 0.0]))


class CarDynamics(Dynamics):
    """Car-like vehicle, state [x, y, h, v, w], effort [force, torque]
    (reference arithmetic: demos/demo_car.py:46-72)."""
    kind = SYS_CAR
    nstates = 5
    ncontrols = 2

    def __init__(self, invM, D, u_lo, u_hi, velmax0):
        self.invM = _f64(invM, (2,))
        self.D = _f64(D, (2,))
        self.u_lo = _f64(u_lo, (2,))
        self.u_hi = _f64(u_hi, (2,))
        self.velmax0 = float(velmax0)

    def blob(self):
        return np.concatenate((self.invM, self.D, self.u_lo, self.u_hi, [self.velmax0]))


class Pendulum2Dynamics(Dynamics):
    """Torque-limited single pendulum, state [theta, omega], effort [tau]
    (synthetic config C1; the reference's demo_pendulum.py
    cannot run).  theta = 0 hangs down."""
    kind = SYS_PENDULUM2
    nstates = 2
    ncontrols = 1

    def __init__(self, mass=1.0, length=1.0, damping=0.1, gravity=9.81, tau_max=np.inf):
        self.mass = float(mass); self.length = float(length)
        self.damping = float(damping); self.gravity = float(gravity)
        self.tau_max = float(tau_max)

    def blob(self):
        return np.array([self.mass, self.length, self.damping, self.gravity, self.tau_max])


class DoubleIntegratorDynamics(Dynamics):
    """D-axis double integrator, state [p (D), v (D)], effort [a (D)]
    (synthetic config C5).  `a_max` saturates the applied
    acceleration inside the dynamics, the way the demos saturate their actuators
    (demo_car.py:57); inf disables it."""
    kind = SYS_DOUBLE_INTEGRATOR

    def __init__(self, axes=6, a_max=np.inf):
        axes = int(axes)
        if not 1 <= axes <= MAX_M:
            raise ValueError("axes must be in 1..{}".format(MAX_M))
        self.axes = axes
        self.a_max = float(a_max)
        self.nstates = 2 * axes
        self.ncontrols = axes

    def blob(self):
        return np.array([float(self.axes), self.a_max])


# =============================================================================
# lqr providers
# =============================================================================

class Lqr(_Spec):
    kind = 0

    def blob(self, n, m):
        raise NotImplementedError

    def __call__(self, x, u):
        """(S, K) = lqr(x, u), evaluated on the device in float64."""
        return self._evaluator().lqr(x, u)


class PDRotatedLqr(Lqr):
    """Hand-tuned PD gain rotated into the body frame with a constant metric:
    boats  K = [kp R(h)^T, kd]           (demos/demo_boat_novice.py:86-103)
    car    K = [kp [[c, s, 0],[0,0,1]], kd]   (demos/demo_car.py:95-113)
    S is any constant n x n matrix (identity in the standalone demos,
    diag(1,1,1,0,0,0) in demos/lqrrt_ros/behaviors/car.py:65)."""
    kind = LQR_PD_ROTATED

    def __init__(self, kp, kd, S=None):
        self.kp = _f64(kp); self.kd = _f64(kd)
        if self.kp.ndim != 2 or self.kp.shape != self.kd.shape or self.kp.shape[0] != self.kp.shape[1]:
            raise ValueError("kp and kd must be square matrices of equal size")
        self.S = None if S is None else _f64(S)

    def blob(self, n, m):
        if self.kp.shape != (m, m):
            raise ValueError("kp/kd must be {0}x{0}".format(m))
        S = np.eye(n) if self.S is None else _f64(self.S, (n, n))
        return np.concatenate((self.kp.ravel(), self.kd.ravel(), S.ravel()))


class ConstantLqr(Lqr):
    """(S, K) independent of (x, u) (demos/demo_pendulum.py:119-126)."""
    kind = LQR_CONSTANT

    def __init__(self, S, K):
        self.S = _f64(S); self.K = _f64(K)

    def blob(self, n, m):
        return np.concatenate((_f64(self.S, (n, n)).ravel(), _f64(self.K, (m, n)).ravel()))


class FDRiccatiLqr(Lqr):
    """Built-in provider named by BASELINE.json's north_star (not present in
    the reference, SURVEY.md M2 -- parity unpinned): central-difference
    Jacobians of the device dynamics about (x, u), then `steps` backward
    iterations of the discrete Riccati recursion from P = Q;  S = P_0, K = K_0
    with  u = K erf(xtar, x)  (sign convention of planner.py:386-387).
    `snap` > 0 rounds every Jacobian entry to the nearest multiple of it before the
    sweep: for (piecewise) linear dynamics whose entries lie on that grid this removes
    the finite-difference noise (float32 differences at |x| ~ 50 are noisy at the 1e-3
    level) and lets a rollout keep its gain while the Jacobian does not change."""
    kind = LQR_FD_RICCATI

    def __init__(self, Q, R, steps=20, eps=1e-3, snap=0.0):
        self.Q = _f64(Q); self.R = _f64(R)
        self.steps = int(steps); self.eps = float(eps); self.snap = float(snap)
        if self.steps < 1:
            raise ValueError("steps must be >= 1")
        if self.snap < 0:
            raise ValueError("snap must be >= 0")

    def blob(self, n, m):
        return np.concatenate((_f64(self.Q, (n, n)).ravel(), _f64(self.R, (m, m)).ravel(),
                               [self.eps, float(self.steps), self.snap]))


# =============================================================================
# error function
# =============================================================================

class Erf(_Spec):
    """e = xgoal - x with the listed dims wrapped onto (-pi, pi]
    (demos/demo_boat_novice.py:106-118; default np.subtract, planner.py:87)."""

    def __init__(self, angle_dims=()):
        self.angle_dims = tuple(int(d) for d in angle_dims)

    def mask(self, n):
        mask = 0
        for d in self.angle_dims:
            if not 0 <= d < n:
                raise ValueError("angle dim {} out of range".format(d))
            mask |= 1 << d
        return mask

    def __call__(self, xgoal, x):
        return self._evaluator().erf(xgoal, x)


# =============================================================================
# feasibility primitives; a Feasibility is their conjunction
# =============================================================================

class _Prim(object):
    kind = 0
    dims = (0, 0, 0, 0)
    flags = 0
    inflate = 0.0
    obs = None      # (O, stride) float64
    verts = None    # (2, V) float64
    lo = None
    hi = None


class StateBox(_Prim):
    """Infeasible if x[i] > hi[i] or x[i] < lo[i] for any i (body-velocity
    limits of demos/demo_boat_advanced.py:205-213); +-inf disables a bound."""
    kind = FEAS_STATE_BOX

    def __init__(self, lo, hi):
        self.lo = _f64(lo); self.hi = _f64(hi)


class EffortBox(_Prim):
    """Infeasible if |u[i]| > umax[i] (demos/demo_pendulum.py:154-157)."""
    kind = FEAS_EFFORT_BOX

    def __init__(self, umax):
        self.hi = _f64(np.atleast_1d(umax)); self.lo = -self.hi


class StateRegion(_Prim):
    """Forbidden axis-aligned region of state space: infeasible if
    lo[i] <= x[i] <= hi[i] for every i (config C1's obstacle region)."""
    kind = FEAS_STATE_REGION

    def __init__(self, lo, hi):
        self.lo = _f64(lo); self.hi = _f64(hi)


def _live(obs, stride, extra=0.0):
    """Drop obstacles that can never be hit: the reference marks inactive grid
    slots with [-9999, -9999, -9999] (demo_boat_novice.py:151) and tests
    `norm <= r (+ inflate)`, which a negative bound never satisfies."""
    obs = np.array(obs, dtype=np.float64).reshape(-1, stride)
    return np.ascontiguousarray(obs[(obs[:, stride - 1] + extra) >= 0])


class CirclesPoint(_Prim):
    """Point robot vs inflated circles [cx, cy, r]: infeasible if
    ||x[dims] - c|| <= inflate + r (demos/demo_boat_novice.py:160-164)."""
    kind = FEAS_CIRCLES_POINT

    def __init__(self, obs, inflate=0.0, dims=(0, 1)):
        self.inflate = float(inflate)
        self.obs = _live(obs, 3, self.inflate)
        self.dims = (int(dims[0]), int(dims[1]), 0, 0)


class CirclesCloud(_Prim):
    """Body-frame vertex cloud `verts` (2 x V) rotated by heading and translated
    vs circles: infeasible if any vertex is <= r from any centre
    (demos/demo_boat_intermediate.py:198-210).  double_center reproduces
    demos/demo_car.py:175, whose extra row makes a vertex at 2*x[:2]."""
    kind = FEAS_CIRCLES_CLOUD

    def __init__(self, obs, verts, dims=(0, 1, 2), double_center=False):
        self.obs = _live(obs, 3)
        self.verts = _f64(verts)
        if self.verts.ndim != 2 or self.verts.shape[0] != 2:
            raise ValueError("verts must be 2 x V")
        self.dims = (int(dims[0]), int(dims[1]), int(dims[2]), 0)
        self.flags = 1 if double_center else 0


class BoxesCloud(_Prim):
    """Vertex cloud vs axis-aligned boxes [cx, cy, hx, hy]: infeasible if any
    vertex lies inside (closed) any box.  New primitive for BASELINE.json's
    '20 box obstacles' (the reference has circles only)."""
    kind = FEAS_BOXES_CLOUD

    def __init__(self, boxes, verts, dims=(0, 1, 2), double_center=False):
        self.obs = np.ascontiguousarray(np.array(boxes, dtype=np.float64).reshape(-1, 4))
        self.verts = _f64(verts)
        if self.verts.ndim != 2 or self.verts.shape[0] != 2:
            raise ValueError("verts must be 2 x V")
        self.dims = (int(dims[0]), int(dims[1]), int(dims[2]), 0)
        self.flags = 1 if double_center else 0


class SpheresPoint(_Prim):
    """Point vs inflated spheres [cx, cy, cz, r] over three state dims (C5)."""
    kind = FEAS_SPHERES_POINT

    def __init__(self, obs, inflate=0.0, dims=(0, 1, 2)):
        self.inflate = float(inflate)
        self.obs = _live(obs, 4, self.inflate)
        self.dims = (int(dims[0]), int(dims[1]), int(dims[2]), 0)


class OccupancyGrid(_Prim):
    """Hull vertex cloud vs an occupancy grid, the ROS node's is_feasible
    (demos/lqrrt_ros/nodes/lqrrt_node.py:719-745): every vertex is mapped to a cell by truncating
    cpm * (vertex - origin) toward zero; infeasible if any cell value is >= threshold or any index is
    past the grid's end.  Negative indices wrap around as NumPy's do (and are out of range below -size),
    which is what the reference does, deliberately or not."""
    kind = FEAS_OGRID_CLOUD

    def __init__(self, grid, cpm, origin, threshold, verts, dims=(0, 1, 2)):
        g = np.ascontiguousarray(np.array(grid, dtype=np.float64))
        if g.ndim != 2:
            raise ValueError("grid must be rows x cols")
        self.grid_shape = g.shape
        self.obs = g.reshape(g.shape[0], g.shape[1])
        self.verts = _f64(verts)
        if self.verts.ndim != 2 or self.verts.shape[0] != 2:
            raise ValueError("verts must be 2 x V")
        self.cpm = float(cpm); self.origin = _f64(origin, (2,)); self.threshold = float(threshold)
        self.dims = (int(dims[0]), int(dims[1]), int(dims[2]), 0)
        self.lo = np.array([float(g.shape[1]), self.cpm, self.origin[0], self.origin[1], self.threshold])
        self.hi = np.zeros(5)


class Feasibility(_Spec):
    """is_feasible(x, u) = all primitives pass, evaluated in list order."""

    def __init__(self, prims=()):
        self.prims = list(prims)
        for p in self.prims:
            if not isinstance(p, _Prim):
                raise ValueError("Feasibility takes a list of primitives")

    def __call__(self, x, u):
        return self._evaluator().is_feasible(x, u)


# =============================================================================
# bundle
# =============================================================================

class System(object):
    """The four callbacks of one vehicle, bundled so that each can find the
    others (FDRiccatiLqr differentiates the dynamics; a bare callback call needs
    a device evaluator).  Pass ``sys.dynamics, sys.lqr, sys.erf,
    sys.is_feasible`` where the reference takes its four functions."""

    def __init__(self, dynamics, lqr, erf=None, is_feasible=None, dt_hint=0.1, name=''):
        if not isinstance(dynamics, Dynamics):
            raise ValueError("dynamics must be a lqrrt_b200.models.Dynamics spec")
        if not isinstance(lqr, Lqr):
            raise ValueError("lqr must be a lqrrt_b200.models.Lqr spec")
        self.dynamics = dynamics
        self.lqr = lqr
        self.erf = Erf() if erf is None else erf
        self.is_feasible = Feasibility() if is_feasible is None else is_feasible
        self.nstates = dynamics.nstates
        self.ncontrols = dynamics.ncontrols
        self.dt_hint = float(dt_hint)
        self.name = name
        self._eval = None
        for s in (self.dynamics, self.lqr, self.erf, self.is_feasible):
            s._bundle = self

    def evaluator(self):
        if self._eval is None:
            from ._native import ModelEvaluator
            self._eval = ModelEvaluator(self)
        return self._eval


def system_digest(system):
    """Content key of everything a System uploads to the device (dynamics / lqr blobs, erf mask, every
    feasibility table).  The Planner rebuilds its engine when this changes, so a spec edited in place between
    plans (new obstacles in Feasibility.prims[i].obs, retuned gains) is never planned against stale tables."""
    n, m = system.nstates, system.ncontrols
    parts = [('sys', system.dynamics.kind), np.asarray(system.dynamics.blob(), dtype=np.float64).tobytes(),
             ('lqr', system.lqr.kind), np.asarray(system.lqr.blob(n, m), dtype=np.float64).tobytes(),
             ('erf', system.erf.mask(n))]
    for p in system.is_feasible.prims:
        parts.append((p.kind, p.flags, float(p.inflate), tuple(int(d) for d in p.dims)))
        for a in (p.obs, p.verts, p.lo, p.hi):
            parts.append(b'' if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
    return hash(tuple(parts))


# verified Systems of plain-callable models: id-tuple -> (the callables themselves, fingerprint of the globals
# they close over, System).  Holding the callables keeps their ids from being reused by reloaded code, and the
# fingerprint notices a user who rebinds or mutates `obs`, `vps`, the gains ... between plans -- the model is
# then recognised (and verified against the user's functions) again instead of planning on stale tables.
_RECOGNISED = {}
_GLOBAL_NAMES = ('B', 'D', 'D_neg', 'D_pos', 'invB', 'invM', 'kd', 'kp', 'magic_rudder', 'thrust_max', 'u_max',
                 'velmax', 'velmax_neg', 'velmax_neg_plan', 'velmax_pos', 'velmax_pos_plan', 'vps', 'obs',
                 'boat_length', 'car_length', 'planning', 'umax_plan')


def _globals_fingerprint(fns):
    import hashlib
    h = hashlib.sha1()
    for fn in fns:
        g = getattr(fn, '__globals__', None)
        if g is None:
            continue
        for name in _GLOBAL_NAMES:
            if name in g:
                h.update(name.encode())
                try:
                    h.update(np.ascontiguousarray(np.array(g[name], dtype=np.float64)).tobytes())
                except Exception:
                    h.update(repr(g[name]).encode())
    return h.hexdigest()


def bundle_of(dynamics, lqr, erf, is_feasible, nstates=None, ncontrols=None):
    """Resolve the four Planner arguments to one System, or raise.

    Device specs are used as they are.  Plain Python callables (the unchanged reference demos) go through
    `recognize.recognise`: parameters are read from the functions' globals and the resulting device model
    is accepted only if it reproduces the user's own functions on a few hundred probes.  There is no host
    fallback: anything that cannot be verified on the device is an error."""
    specs = (dynamics, lqr, is_feasible)
    if not any(isinstance(s, _Spec) for s in specs) and all(hasattr(s, '__call__') for s in specs) \
            and nstates is not None:
        key = (id(dynamics), id(lqr), id(erf), id(is_feasible))
        fns = (dynamics, lqr, erf, is_feasible)
        fp = _globals_fingerprint(fns)
        hit = _RECOGNISED.get(key)
        if hit is None or hit[1] != fp or any(a is not b for a, b in zip(hit[0], fns)):
            from . import recognize
            from ._native import NativeError
            try:
                _RECOGNISED[key] = (fns, fp, recognize.recognise(dynamics, lqr, erf, is_feasible, int(nstates), int(ncontrols)))
            except recognize.NotRecognised as e:
                raise TypeError(
                    "lqrrt_b200 runs the planner on the GPU and needs device model specs (lqrrt_b200.models); the "
                    "plain Python callables given could not be matched to a compiled model: {}. There is no CPU "
                    "fallback.".format(e))
            except NativeError as e:
                raise TypeError(
                    "plain Python callables can only be accepted after the device model has been verified "
                    "against them on the GPU, and the engine is not available ({}). There is no CPU fallback."
                    .format(e))
        return _RECOGNISED[key][2]
    for s, nm in zip(specs, ('dynamics', 'lqr', 'is_feasible')):
        if not isinstance(s, _Spec):
            raise TypeError(
                "lqrrt_b200 runs the planner on the GPU and needs device model specs; "
                "`{}` is a plain Python callable ({!r}). Build it from lqrrt_b200.models "
                "(e.g. models.boat_novice(obs)) -- there is no CPU fallback.".format(nm, s))
    if erf is np.subtract or erf is None:
        erf_spec = Erf()
    elif isinstance(erf, Erf):
        erf_spec = erf
    else:
        raise TypeError("`erf` must be np.subtract or a lqrrt_b200.models.Erf spec")
    b = dynamics._bundle
    if b is not None and b.lqr is lqr and b.is_feasible is is_feasible and \
            (b.erf is erf_spec or (erf_spec.angle_dims == b.erf.angle_dims)):
        return b
    return System(dynamics, lqr, erf_spec, is_feasible)


# =============================================================================
# the reference's demo vehicles and the synthetic configs as ready-made bundles
# =============================================================================

def vertex_cloud(length, width, buffer, spacing):
    """Body-frame grid of points covering the hull (demo_car.py:77-90)."""
    gx, gy = np.mgrid[slice(-(length + buffer) / 2, (length + buffer) / 2 + spacing, spacing),
                      slice(-(width + buffer) / 2, (width + buffer) / 2 + spacing, spacing)]
    return np.ascontiguousarray(np.vstack((gx.reshape(-1), gy.reshape(-1))))


BOAT_LENGTH = 210 * 0.0254
BOAT_WIDTH = 96 * 0.0254


def boat_novice(obs=()):
    """demos/demo_boat_novice.py:23-38,86-118,160-164."""
    velmax_pos = np.array([1.1, 0.45, 0.2]); velmax_neg = np.array([0.68, 0.45, 0.2])
    thrust_max = 220; thrust_lever = 2.15
    u_max = np.array([2 * np.sqrt(2) * thrust_max, 2 * np.sqrt(2) * thrust_max, 4 * thrust_lever * thrust_max])
    dyn = BoatDynamics(invM=[1 / 500, 1 / 500, 1 / 500], D_pos=np.abs(u_max / velmax_pos),
                       D_neg=np.abs(u_max / velmax_neg), u_max=u_max, clip='wrench')
    lqr = PDRotatedLqr(np.diag([120, 120, 350]), np.diag([120, 120, 100]))
    feas = Feasibility([CirclesPoint(obs, inflate=BOAT_LENGTH / 2)])
    return System(dyn, lqr, Erf([2]), feas, name='boat_novice')


def boat_intermediate(obs=()):
    """demos/demo_boat_intermediate.py:26-41,48-100,103-141,198-210."""
    velmax_pos = np.array([1.1, 0.45, 0.2]); velmax_neg = np.array([0.68, 0.45, 0.2])
    thrust_max = 220; thrust_lever = 2.15
    u_max = np.array([2 * np.sqrt(2) * thrust_max, 0.2 * np.sqrt(2) * thrust_max, 4 * thrust_lever * thrust_max])
    dyn = BoatDynamics(invM=[1 / 500, 1 / 500, 1 / 500], D_pos=np.abs(u_max / velmax_pos),
                       D_neg=np.abs(u_max / velmax_neg), u_max=u_max, clip='wrench',
                       rudder=5000, carlike=True, velmax_pos0=velmax_pos[0], velmax_neg0=velmax_neg[0])
    lqr = PDRotatedLqr(np.diag([120, 120, 0]), np.diag([120, 120, 0]))
    vps = vertex_cloud(BOAT_LENGTH, BOAT_WIDTH, 2, 0.5)
    feas = Feasibility([CirclesCloud(obs, vps)])
    return System(dyn, lqr, Erf([2]), feas, name='boat_intermediate')


def boat_advanced_params():
    velmax_pos = np.array([2.5, 1, 0.7]); velmax_neg = np.array([-0.8, -1, -0.7])
    thrust_max = np.array([220, 220, 220, 220])
    pos = np.array([[-1.9, 1.0, -0.0123], [-1.9, -1.0, -0.0123], [1.6, 0.6, -0.0123], [1.6, -0.6, -0.0123]])
    dirs = np.array([[0.7071, 0.7071, 0.0], [0.7071, -0.7071, 0.0], [0.7071, -0.7071, 0.0], [0.7071, 0.7071, 0.0]])
    levers = np.cross(pos, dirs)
    B = np.concatenate((dirs.T, levers.T))[[0, 1, 5]]
    invB = np.linalg.pinv(B)
    Fx = B.dot(thrust_max * [1, 1, 1, 1])[0]
    Fy = B.dot(thrust_max * [1, -1, -1, 1])[1]
    Mz = B.dot(thrust_max * [-1, 1, -1, 1])[2]
    D_pos = np.abs([Fx, Fy, Mz] / velmax_pos)
    D_neg = np.abs([Fx, Fy, Mz] / velmax_neg)
    return dict(velmax_pos=velmax_pos, velmax_neg=velmax_neg, thrust_max=thrust_max, B=B, invB=invB,
                D_pos=D_pos, D_neg=D_neg)


def boat_advanced(obs=()):
    """demos/demo_boat_advanced.py:25-68,75-165,205-225."""
    p = boat_advanced_params()
    dyn = BoatDynamics(invM=[1 / 500, 1 / 500, 1 / 500], D_pos=p['D_pos'], D_neg=p['D_neg'],
                       B=p['B'], invB=p['invB'], thrust_max=p['thrust_max'], clip='thruster',
                       rudder=4000, carlike=True, velmax_pos0=p['velmax_pos'][0], velmax_neg0=p['velmax_neg'][0])
    lqr = PDRotatedLqr(np.diag([120, 20, 0]), np.diag([120, 20, 0]))
    vps = vertex_cloud(BOAT_LENGTH, BOAT_WIDTH, 0.25, 1)
    lo = np.array([-np.inf, -np.inf, -np.inf, -0.65, -0.4, -0.2])
    hi = np.array([np.inf, np.inf, np.inf, 1.1, 0.4, 0.2])
    feas = Feasibility([StateBox(lo, hi), CirclesCloud(obs, vps)])
    return System(dyn, lqr, Erf([2]), feas, name='boat_advanced')


def car(obs=(), boxes=None):
    """demos/demo_car.py:28-39,46-126,168-180; `boxes` adds BASELINE.json's
    box-obstacle variant (config C2) on top of / instead of the circles."""
    velmax = np.array([1.1, 1.0]); u_max = np.array([650.0, 1800.0])
    dyn = CarDynamics(invM=[1 / 500, 1 / 500], D=np.abs(u_max / velmax),
                      u_lo=[-u_max[0] / 10, -u_max[1]], u_hi=u_max, velmax0=velmax[0])
    lqr = PDRotatedLqr(np.diag([120, 600]), np.diag([120, 600]))
    vps = vertex_cloud(6, 3, 2, 0.5)
    prims = []
    if len(obs):
        prims.append(CirclesCloud(obs, vps, double_center=True))
    if boxes is not None and len(boxes):
        prims.append(BoxesCloud(boxes, vps, double_center=True))
    return System(dyn, lqr, Erf([2]), Feasibility(prims), name='car')


def pendulum2(tau_max=5.0, region=((1.2, -1.0), (1.8, 1.0)), Q=(10.0, 1.0), R=(1.0,), steps=50, eps=1e-3):
    """Config C1: torque-limited swing-up with one forbidden
    state-space box and the reference pendulum's effort constraint."""
    dyn = Pendulum2Dynamics(1.0, 1.0, 0.1, 9.81, tau_max=np.inf)
    lqr = FDRiccatiLqr(np.diag(Q), np.diag(R), steps=steps, eps=eps)
    prims = [EffortBox([tau_max])]
    if region is not None:
        prims.append(StateRegion(region[0], region[1]))
    return System(dyn, lqr, Erf([0]), Feasibility(prims), dt_hint=0.01, name='pendulum2')


def double_integrator(axes=6, spheres=(), a_max=1.0, steps=20, eps=1e-3, riccati=True, v_max=3.0, snap=0.0):
    """Config C5: accelerations saturate at a_max inside the dynamics (SURVEY.md's
    "|u| <= 1 via is_feasible" would reject every rollout whose first LQR effort exceeds 1, i.e. nearly all
    of them, so the limit is applied where the demos apply theirs); speeds beyond v_max are infeasible."""
    dyn = DoubleIntegratorDynamics(axes, a_max=a_max)
    n, m = dyn.nstates, dyn.ncontrols
    if riccati:
        # the true Jacobian entries are 1, dt and 0 (and anything between 0 and dt where a difference straddles the
        # saturation): with snap = dt / 4 (what configs.c5_double_integrator passes) snapping restores them from the
        # noisy differences
        lqr = FDRiccatiLqr(np.eye(n), np.eye(m), steps=steps, eps=eps, snap=snap)
    else:
        K = np.hstack((np.eye(m), 1.7 * np.eye(m)))
        lqr = ConstantLqr(np.eye(n), K)
    prims = [StateBox([-np.inf] * m + [-v_max] * m, [np.inf] * m + [v_max] * m)]
    if len(spheres):
        prims.append(SpheresPoint(spheres, dims=(0, 1, 2)))
    return System(dyn, lqr, Erf(), Feasibility(prims), name='double_integrator{}'.format(n))
