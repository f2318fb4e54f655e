"""A user's planar-boat model written the way the reference demos write theirs: plain functions over module
globals with the demos' parameter names (cf. demos/demo_boat_advanced.py:25-225).  Different numbers from
any demo, so that recognition has to read them rather than match a preset.  Used by
tests/test_gpu_recognize.py to check that unchanged closure-style models are accepted after on-device
verification."""
import numpy as np

nstates, ncontrols = 6, 3
invM = np.array([1 / 420.0, 1 / 420.0, 1 / 610.0])
velmax_pos = np.array([2.0, 0.8, 0.6])
velmax_neg = np.array([-0.7, -0.8, -0.6])
thrust_max = np.array([180.0, 180.0, 200.0, 200.0])
_pos = np.array([[-1.7, 0.9, 0.0], [-1.7, -0.9, 0.0], [1.5, 0.5, 0.0], [1.5, -0.5, 0.0]])
_dir = np.array([[0.8, 0.6, 0.0], [0.8, -0.6, 0.0], [0.6, -0.8, 0.0], [0.6, 0.8, 0.0]])
B = np.concatenate((_dir.T, np.cross(_pos, _dir).T))[[0, 1, 5]]
invB = np.linalg.pinv(B)
D_pos = np.abs(np.array([500.0, 420.0, 900.0]) / velmax_pos)
D_neg = np.abs(np.array([500.0, 420.0, 900.0]) / velmax_neg)
magic_rudder = 2500.0
kp = np.diag([90.0, 30.0, 0.0])
kd = np.diag([110.0, 25.0, 0.0])
boat_length, boat_width = 5.0, 2.2
gx, gy = np.mgrid[-2.75:3.0:1.1, -1.25:1.5:0.625]
vps = np.vstack((gx.reshape(-1), gy.reshape(-1)))
_rng = np.random.RandomState(3)
obs = np.column_stack((_rng.uniform(8, 55, 24), _rng.uniform(8, 55, 24), _rng.uniform(0.8, 2.5, 24)))
velmax_pos_plan = np.array([1.0, 0.35, 0.25])
velmax_neg_plan = np.array([-0.5, -0.35, -0.25])


def dynamics(x, u, dt):
    c, s = np.cos(x[2]), np.sin(x[2])
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    D = np.where(x[3:] >= 0, D_pos, D_neg)
    vw = R[:2, :2].dot(x[3:5])
    ang = np.arctan2(vw[1], vw[0])
    u[2] = u[2] + magic_rudder * np.arctan2(np.sin(ang) * c - np.cos(ang) * s, np.cos(ang) * c + np.sin(ang) * s)
    u = B.dot(np.clip(invB.dot(u), -thrust_max, thrust_max))
    xn = x + np.concatenate((R.dot(x[3:]), invM * (u - D * x[3:]))) * dt
    if x[3] > 0:
        xn[5] = np.clip(np.abs(xn[3] / velmax_pos[0]), 0, 1) * xn[5]
    elif x[3] < 0:
        xn[5] = np.clip(np.abs(xn[3] / velmax_neg[0]), 0, 1) * xn[5]
    if xn[3] < 0:
        xn[3] = 0
    return xn


def lqr(x, u):
    c, s = np.cos(x[2]), np.sin(x[2])
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    return (np.eye(6), np.hstack((kp.dot(R.T), kd)))


def erf(xgoal, x):
    e = xgoal - x
    e[2] = np.arctan2(np.sin(xgoal[2] - x[2]), np.cos(xgoal[2] - x[2]))
    return e


def is_feasible(x, u):
    if np.any(x[3:] > velmax_pos_plan) or np.any(x[3:] < velmax_neg_plan):
        return False
    c, s = np.cos(x[2]), np.sin(x[2])
    verts = x[:2] + np.array([[c, -s], [s, c]]).dot(vps).T
    for ob in obs:
        if np.any(np.linalg.norm(verts - ob[:2], axis=1) <= ob[2]):
            return False
    return True
