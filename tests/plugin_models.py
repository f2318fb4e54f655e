"""User-dynamics sources shared by the CPU build test and the GPU parity test (lqrrt_b200/plugin.py)."""
import numpy as np

from lqrrt_b200 import configs, models, plugin

DINT2_STEP = '''
for (int i = 0; i < 2; i++) {
    u[i] = min_t(max_t(u[i], -p[0]), p[0]);
    xn[i] = x[i] + x[2 + i] * dt;
    xn[2 + i] = x[2 + i] + u[i] * dt;
}'''

# planar unicycle with first-order speed / turn-rate lags: state [x, y, h, v, w], effort [a, alpha]
UNICYCLE_STEP = '''
const T a = min_t(max_t(u[0], -p[0]), p[0]);
const T al = min_t(max_t(u[1], -p[1]), p[1]);
xn[0] = x[0] + x[3] * c * dt;
xn[1] = x[1] + x[3] * s * dt;
xn[2] = x[2] + x[4] * dt;
xn[3] = x[3] + (a - p[2] * x[3]) * dt;
xn[4] = x[4] + (al - p[3] * x[4]) * dt;
u[0] = a; u[1] = al;'''
UNICYCLE_PARAMS = [2.0, 3.0, 0.4, 1.5]


def unicycle_numpy(x, u, dt, p=UNICYCLE_PARAMS):
    a = min(max(u[0], -p[0]), p[0]); al = min(max(u[1], -p[1]), p[1])
    return np.array([x[0] + x[3] * np.cos(x[2]) * dt, x[1] + x[3] * np.sin(x[2]) * dt, x[2] + x[4] * dt,
                     x[3] + (a - p[2] * x[3]) * dt, x[4] + (al - p[3] * x[4]) * dt])


def dint2_pair(riccati, seed=0, spheres=60):
    """(problem with the built-in 2-axis double integrator, the same problem with user-compiled dynamics)."""
    base = configs.c5_double_integrator(seed, count=0, axes=2, max_nodes=4000, riccati=riccati)
    b = base.system
    dyn = plugin.UserDynamics(4, 2, DINT2_STEP, params=[b.dynamics.a_max], name='dint2')
    lqr = models.FDRiccatiLqr(b.lqr.Q, b.lqr.R, steps=b.lqr.steps, eps=b.lqr.eps, snap=b.lqr.snap) if riccati \
        else models.ConstantLqr(b.lqr.S, b.lqr.K)
    rng = np.random.RandomState(seed)
    circles = np.column_stack((5 + 90 * rng.rand(spheres, 2), 1 + 2 * rng.rand(spheres)))
    prims_a = [models.StateBox(b.is_feasible.prims[0].lo, b.is_feasible.prims[0].hi), models.CirclesPoint(circles)]
    prims_b = [models.StateBox(b.is_feasible.prims[0].lo, b.is_feasible.prims[0].hi), models.CirclesPoint(circles)]
    sys_a = models.System(b.dynamics, b.lqr, models.Erf(), models.Feasibility(prims_a), name='dint2-builtin')
    sys_b = models.System(dyn, lqr, models.Erf(), models.Feasibility(prims_b), name='dint2-user')

    def prob(system):
        return configs.Problem(system.name, system, base.x0, base.goal, base.goal_buffer, base.error_tol,
                               base.sample_space, base.goal_bias, base.horizon, base.dt, base.FPR,
                               max_nodes=4000, min_time=60, max_time=60)
    return prob(sys_a), prob(sys_b)


def unicycle_problem():
    dyn = plugin.UserDynamics(5, 2, UNICYCLE_STEP, params=UNICYCLE_PARAMS, heading=2, name='unicycle')
    lqr = models.FDRiccatiLqr(np.diag([1.0, 1.0, 1.0, 0.1, 0.1]), np.eye(2), steps=15, eps=1e-4)
    rng = np.random.RandomState(4)
    circles = np.column_stack((5 + 40 * rng.rand(25, 2), 1 + rng.rand(25)))
    system = models.System(dyn, lqr, models.Erf([2]), models.Feasibility([models.CirclesPoint(circles, inflate=0.5)]),
                           name='unicycle')
    gb = [1.5, 1.5, np.inf, np.inf, np.inf]
    ss = [(0, 50), (0, 50), (-np.pi, np.pi), (0, 3), (-1, 1)]
    return configs.Problem('unicycle', system, np.zeros(5), np.array([45.0, 45.0, 0, 0, 0]), gb, [0.5, 0.5, 0.3, 1, 1], ss,
                           [0.1, 0.1, 0, 0, 0], horizon=2, dt=0.1, FPR=0.5, max_nodes=3000, min_time=60, max_time=60)


# every (model, dtype, lqr provider) the GPU tests use; the CPU test builds them so the libraries travel
BUILDS = [('dint2', 'f64', models.LQR_FD_RICCATI), ('dint2', 'f64', models.LQR_CONSTANT),
          ('dint2', 'f32', models.LQR_CONSTANT), ('unicycle', 'f64', models.LQR_FD_RICCATI)]


def dynamics_of(name):
    return dint2_pair(True)[1].system.dynamics if name == 'dint2' else unicycle_problem().system.dynamics
