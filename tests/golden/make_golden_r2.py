"""
Round-2 additions to tests/golden/, generated from the UNMODIFIED reference at /root/reference like
make_golden.py (SURVEY.md section 8c recipe; build container only):

    python tests/golden/make_golden_r2.py        ->  tests/golden/r2_extra.npz

  fa_*   `Planner._steer(ID, xtar, force_arrive=True)` (planner.py:294-303, 399-410) on the boat_novice tree of
         make_golden.py, for targets it reaches (np.allclose exit) and targets behind an obstacle (infeasible exit
         with the FPR cut); the wall-clock limit never triggers (min_time / 2 = 200 s).
  ds_*   `Planner._costs_to_go` minima / argmins / pruned picks on the demo_car tree with the ROS car's metric
         S = diag(1, 1, 1, 0, 0) (demos/lqrrt_ros/behaviors/car.py:65) instead of the demo's identity.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

K = 400


def force_arrive(out):
    planner, ns = ref_shim.run_reference_demo('demo_boat_novice', 0, K, dict(horizon=2, dt=0.1, FPR=0.5))
    tree = planner.tree
    obs = np.array(ns['obs'], dtype=np.float64)
    live = obs[obs[:, 0] > -1000]
    rng = np.random.RandomState(7)
    ids = []; tars = []
    # (a) reachable: a pose 1-3 m from the node, at rest, in free water
    tries = 0
    while len(ids) < 4 and tries < 1000:
        tries += 1
        ID = int(rng.randint(0, tree.size))
        x = tree.state[ID]
        t = np.array([x[0] + 3 * (rng.rand() - 0.5), x[1] + 3 * (rng.rand() - 0.5), x[2] + 0.6 * (rng.rand() - 0.5), 0, 0, 0])
        if np.min(np.linalg.norm(live[:, :2] - t[:2], axis=1)) > 6.0 and np.min(np.linalg.norm(live[:, :2] - x[:2], axis=1)) > 6.0:
            ids.append(ID); tars.append(t)
    # (b) blocked: the target is the mirror image of the node through an obstacle centre
    tries = 0
    while len(ids) < 7 and tries < 4000:
        tries += 1
        ID = int(rng.randint(0, tree.size))
        x = tree.state[ID]
        d = np.linalg.norm(live[:, :2] - x[:2], axis=1)
        o = live[int(np.argmin(d))]
        if 4.5 < d.min() < 9.0:
            t = np.array([2 * o[0] - x[0], 2 * o[1] - x[1], x[2], 0, 0, 0])
            ids.append(ID); tars.append(t)
    L = []; XS = []; US = []
    for ID, t in zip(ids, tars):
        xs, us = planner._steer(ID, t, force_arrive=True)
        L.append(len(xs)); XS.append(np.array(xs).reshape(-1, 6)); US.append(np.array(us).reshape(-1, 3))
    out['fa_state'] = tree.state
    out['fa_ids'] = np.array(ids); out['fa_targets'] = np.array(tars); out['fa_len'] = np.array(L)
    out['fa_x'] = np.vstack(XS); out['fa_u'] = np.vstack(US)
    print('force_arrive lengths', L)


def diag_metric(out):
    planner, ns = ref_shim.run_reference_demo('demo_car', 0, K, dict(horizon=5, dt=0.1, FPR=0))
    S = np.diag([1.0, 1.0, 1.0, 0.0, 0.0])
    lqr0 = planner.lqr
    planner.lqr = lambda x, u: (S, lqr0(x, u)[1])
    rng = np.random.RandomState(11)
    ss = np.array(ns['sample_space'], dtype=np.float64)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(48, 5)
    costs = np.array([planner._costs_to_go(q) for q in Q])
    out['ds_S'] = S; out['ds_state'] = planner.tree.state; out['ds_q'] = Q
    out['ds_argmin'] = np.argmin(costs, axis=1); out['ds_mincost'] = np.min(costs, axis=1)
    # exact ties are expected (two dims carry no weight): record every minimiser so the test can check
    # "lowest index among the exact minima" without depending on NumPy's argsort
    out['ds_ntied'] = np.array([(c == c.min()).sum() for c in costs])
    print('diag metric: ties per query', out['ds_ntied'].tolist()[:12])


if __name__ == '__main__':
    out = {}
    force_arrive(out)
    diag_metric(out)
    path = os.path.join(HERE, 'r2_extra.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), 'bytes')
