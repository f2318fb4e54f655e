"""
Planning problems as data: the reference demos' objectives/heuristics sections
(e.g. demos/demo_boat_novice.py:120-186) and the synthetic scale-ups named by
BASELINE.json `configs` (concrete definitions: SURVEY.md section 8d, C1-C5).

Every function returns a `Problem`: the device `System`, the `Planner` /
`Constraints` arguments, and the `update_plan` arguments.  No model arithmetic
here -- only parameters and seeded obstacle-field generators.
"""
from __future__ import annotations

import numpy as np

from . import models


class Problem(object):
    def __init__(self, name, system, x0, goal, goal_buffer, error_tol, sample_space, goal_bias,
                 horizon, dt, FPR, max_nodes=1E5, min_time=1.0, max_time=2.0, obs=None):
        self.name = name
        self.system = system
        self.x0 = np.array(x0, dtype=np.float64)
        self.goal = np.array(goal, dtype=np.float64)
        self.goal_buffer = np.array(goal_buffer, dtype=np.float64)
        self.error_tol = np.array(error_tol, dtype=np.float64)
        self.sample_space = [tuple(map(float, s)) for s in sample_space]
        self.goal_bias = list(goal_bias)
        self.horizon = horizon
        self.dt = dt
        self.FPR = FPR
        self.max_nodes = max_nodes
        self.min_time = min_time
        self.max_time = max_time
        self.obs = obs

    @property
    def nstates(self):
        return self.system.nstates

    @property
    def ncontrols(self):
        return self.system.ncontrols

    def make_planner(self, **kw):
        """Planner + Constraints exactly as a reference demo's PLAN section
        builds them (demo_boat_novice.py:179-186)."""
        from . import Constraints, Planner
        s = self.system
        constraints = Constraints(nstates=s.nstates, ncontrols=s.ncontrols,
                                  goal_buffer=self.goal_buffer, is_feasible=s.is_feasible)
        args = dict(horizon=self.horizon, dt=self.dt, FPR=self.FPR, error_tol=self.error_tol,
                    erf=s.erf, min_time=self.min_time, max_time=self.max_time,
                    max_nodes=self.max_nodes, goal0=self.goal, printing=False)
        args.update(kw)
        return Planner(s.dynamics, s.lqr, constraints, **args)


# ----------------------------------------------------------------------------
# obstacle fields
# ----------------------------------------------------------------------------

def demo_grid_obstacles(rng, x0, goal, keepout, spacing=12, span=(5, 60)):
    """The demos' noised 6x6 grid of r=1 circles (demo_boat_novice.py:143-155),
    drawing from `rng.rand` in the demos' order; slots inside the keep-out of
    the start or goal stay as the reference's [-9999]*3 sentinel."""
    gx, gy = np.mgrid[slice(span[0], span[1] + spacing, spacing), slice(span[0], span[1] + spacing, spacing)]
    gx = gx.reshape(gx.size); gy = gy.reshape(gy.size)
    obs = -9999 * np.ones((gx.size, 3))
    for i in range(len(obs)):
        p = np.round([gx[i], gy[i]] + 3 * (rng.rand(2) - 0.5), 2)
        if np.linalg.norm(p - goal[:2]) > keepout and np.linalg.norm(p - x0[:2]) > keepout:
            obs[i] = [p[0], p[1], 1]
    return obs


def random_circles(rng, count, lo, hi, rmin, rmax, x0, goal, keepout):
    """`count` circles, centres U([lo,hi]^2) rejecting the keep-out of start and
    goal, radii U(rmin, rmax)."""
    out = []
    while len(out) < count:
        p = lo + (hi - lo) * rng.rand(2)
        r = rmin + (rmax - rmin) * rng.rand()
        if np.linalg.norm(p - goal[:2]) > keepout + r and np.linalg.norm(p - x0[:2]) > keepout + r:
            out.append([p[0], p[1], r])
    return np.array(out)


def random_boxes(rng, count, lo, hi, hmin, hmax, x0, goal, keepout):
    out = []
    while len(out) < count:
        p = lo + (hi - lo) * rng.rand(2)
        h = hmin + (hmax - hmin) * rng.rand(2)
        if np.linalg.norm(p - goal[:2]) > keepout + h.max() and np.linalg.norm(p - x0[:2]) > keepout + h.max():
            out.append([p[0], p[1], h[0], h[1]])
    return np.array(out)


def random_spheres(rng, count, lo, hi, rmin, rmax, x0, goal, keepout):
    out = []
    while len(out) < count:
        p = lo + (hi - lo) * rng.rand(3)
        r = rmin + (rmax - rmin) * rng.rand()
        if np.linalg.norm(p - goal[:3]) > keepout + r and np.linalg.norm(p - x0[:3]) > keepout + r:
            out.append([p[0], p[1], p[2], r])
    return np.array(out)


# ----------------------------------------------------------------------------
# the reference's four runnable demos, default obstacle choices
# ----------------------------------------------------------------------------

def demo_boat_novice(seed=0):
    """demos/demo_boat_novice.py:120-186 with obs_choice='grid'."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(6); goal = np.array([40, 40, np.deg2rad(90), 0, 0, 0])
    obs = demo_grid_obstacles(rng, x0, goal, 2 * models.BOAT_LENGTH)
    gb = [6, 6, np.inf, np.inf, np.inf, np.inf]
    ss = [(0, 40), (0, 40), (-np.pi, np.pi), (0.5 * 1.1, 1.1), (-0.45, 0.45), (-0.2, 0.2)]
    return Problem('demo_boat_novice', models.boat_novice(obs), x0, goal, gb, np.array(gb) / 2, ss,
                   [0.5, 0.5, 0, 0, 0, 0], horizon=2, dt=0.1, FPR=0.5, min_time=1, max_time=2, obs=obs)


def demo_boat_intermediate(seed=0):
    """demos/demo_boat_intermediate.py:160-233."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(6); goal = np.array([40, 40, np.deg2rad(90), 0, 0, 0])
    obs = demo_grid_obstacles(rng, x0, goal, 2 * models.BOAT_LENGTH)
    gb = [8, 8, np.inf, np.inf, np.inf, np.inf]
    ss = [(0, 40), (0, 40), (0, 0), (0.9 * 1.1, 1.1), (-0.45, 0.45), (-0.2, 0.2)]
    return Problem('demo_boat_intermediate', models.boat_intermediate(obs), x0, goal, gb, np.array(gb) / 8, ss,
                   [0.2, 0.2, 0, 0, 0, 0], horizon=2, dt=0.1, FPR=0.5, min_time=2, max_time=3, obs=obs)


def demo_boat_advanced(seed=0):
    """demos/demo_boat_advanced.py:168-250."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(6); goal = np.array([40, 40, np.deg2rad(90), 0, 0, 0])
    obs = demo_grid_obstacles(rng, x0, goal, 2 * models.BOAT_LENGTH)
    gb = [8, 8, np.inf, np.inf, np.inf, np.inf]
    ss = [(0, 40), (0, 40), (0, 0), (0.9 * 1.1, 1.1), (-0.4, 0.4), (-0.2, 0.2)]
    return Problem('demo_boat_advanced', models.boat_advanced(obs), x0, goal, gb, np.array(gb) / 8, ss,
                   [0.2, 0.2, 0, 0, 0, 0], horizon=2, dt=0.1, FPR=0.9, min_time=2, max_time=3, obs=obs)


def demo_car(seed=0, obs_choice='some'):
    """demos/demo_car.py:128-206 (default obs_choice='some': three circles)."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(5); goal = np.array([40, 40, np.deg2rad(90), 0, 0])
    if obs_choice == 'some':
        obs = np.array([[20, 20, 5], [10, 30, 2], [40, 10, 3]], dtype=np.float64)
    else:
        obs = demo_grid_obstacles(rng, x0, goal, 2 * 6)
    gb = [8, 8, np.inf, np.inf, np.inf]
    ss = [(0, 80), (0, 80), (-np.pi, np.pi), (0.9 * 1.1, 1.1), (-1, 1)]
    return Problem('demo_car', models.car(obs), x0, goal, gb, np.array(gb) / 2, ss,
                   [0.5, 0.5, 0, 0, 0], horizon=5, dt=0.1, FPR=0, min_time=2, max_time=3, obs=obs)


# ----------------------------------------------------------------------------
# BASELINE.json configs
# ----------------------------------------------------------------------------

def c1_pendulum(seed=0):
    """C1: 2-state torque-limited swing-up, one forbidden state-space region."""
    system = models.pendulum2()
    x0 = [0.0, 0.0]; goal = [np.pi, 0.0]
    return Problem('c1_pendulum2', system, x0, goal, [0.1, 0.3], [0.05, 0.15],
                   [(-np.pi, np.pi), (-8, 8)], [0.1, 0.1], horizon=0.5, dt=0.01, FPR=0.5,
                   min_time=1, max_time=2)


def c2_car(seed=0, boxes=True):
    """C2: the demo car among 20 box obstacles (or 20 circles, boxes=False)."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(5); goal = np.array([40, 40, np.deg2rad(90), 0, 0])
    if boxes:
        bx = random_boxes(rng, 20, 0.0, 80.0, 1.0, 3.0, x0, goal, 2 * 6)
        system = models.car((), boxes=bx); obs = bx
    else:
        obs = random_circles(rng, 20, 0.0, 80.0, 1.0, 3.0, x0, goal, 2 * 6)
        system = models.car(obs)
    gb = [8, 8, np.inf, np.inf, np.inf]
    ss = [(0, 80), (0, 80), (-np.pi, np.pi), (0.9 * 1.1, 1.1), (-1, 1)]
    return Problem('c2_car', system, x0, goal, gb, np.array(gb) / 2, ss, [0.5, 0.5, 0, 0, 0],
                   horizon=5, dt=0.1, FPR=0, min_time=2, max_time=3, obs=obs)


def c3_boat_novice(seed=0, count=50):
    """C3: boat_novice among 50 r=1 circles on [5,60]^2."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(6); goal = np.array([40, 40, np.deg2rad(90), 0, 0, 0])
    obs = random_circles(rng, count, 5.0, 60.0, 1.0, 1.0, x0, goal, 2 * models.BOAT_LENGTH)
    gb = [6, 6, np.inf, np.inf, np.inf, np.inf]
    ss = [(0, 40), (0, 40), (-np.pi, np.pi), (0.5 * 1.1, 1.1), (-0.45, 0.45), (-0.2, 0.2)]
    return Problem('c3_boat_novice', models.boat_novice(obs), x0, goal, gb, np.array(gb) / 2, ss,
                   [0.5, 0.5, 0, 0, 0, 0], horizon=2, dt=0.1, FPR=0.5, obs=obs)


def c4_boat_advanced(seed=0, count=500, arena=200.0, max_nodes=1E5):
    """C4: boat_advanced, 500 circles r~U(0.5,2) on a 200 m arena, goal at the
    far corner, tree grown to 100k nodes -- the scaling benchmark."""
    rng = np.random.RandomState(seed)
    x0 = np.zeros(6); goal = np.array([arena, arena, np.deg2rad(90), 0, 0, 0])
    obs = random_circles(rng, count, 0.0, arena, 0.5, 2.0, x0, goal, 2 * models.BOAT_LENGTH)
    gb = [8, 8, np.inf, np.inf, np.inf, np.inf]
    ss = [(0, arena), (0, arena), (0, 0), (0.9 * 1.1, 1.1), (-0.4, 0.4), (-0.2, 0.2)]
    return Problem('c4_boat_advanced', models.boat_advanced(obs), x0, goal, gb, np.array(gb) / 8, ss,
                   [0.2, 0.2, 0, 0, 0, 0], horizon=2, dt=0.1, FPR=0.9, max_nodes=max_nodes,
                   min_time=60, max_time=60, obs=obs)


def c5_double_integrator(seed=0, count=5000, axes=6, max_nodes=1E6, riccati=True):
    """C5: 12-state double integrator, 5000 spheres in (p1,p2,p3), 1M-node tree."""
    rng = np.random.RandomState(seed)
    n = 2 * axes
    x0 = np.zeros(n); goal = np.concatenate((80.0 * np.ones(axes), np.zeros(axes)))
    spheres = random_spheres(rng, count, 5.0, 95.0, 1.0, 3.0, x0, goal, 4.0) if count else ()
    dt = 0.1
    system = models.double_integrator(axes, spheres, a_max=1.0, riccati=riccati, snap=dt / 4)
    gb = [4.0] * axes + [np.inf] * axes
    ss = [(0, 100)] * axes + [(-2, 2)] * axes
    tol = [2.0] * axes + [np.inf] * axes
    return Problem('c5_double_integrator', system, x0, goal, gb, tol, ss, [0.1] * axes + [0] * axes,
                   horizon=2, dt=dt, FPR=0.5, max_nodes=max_nodes, min_time=60, max_time=60, obs=spheres)


BY_NAME = {
    'demo_boat_novice': demo_boat_novice, 'demo_boat_intermediate': demo_boat_intermediate,
    'demo_boat_advanced': demo_boat_advanced, 'demo_car': demo_car,
    'c1': c1_pendulum, 'c2': c2_car, 'c3': c3_boat_novice, 'c4': c4_boat_advanced,
    'c5': c5_double_integrator,
    'c5k': lambda seed=0: c5_double_integrator(seed, riccati=False),      # C5 with a constant LQR gain
}
