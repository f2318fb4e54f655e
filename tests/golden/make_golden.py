"""
Generates tests/golden/*.npz from the UNMODIFIED reference at /root/reference
(only possible in the build container; the GPU box has no reference, which is
why the vectors are committed).  Recipe: SURVEY.md section 8c.

    python tests/golden/make_golden.py

Per runnable demo (default obstacle choice, np.random.seed(0), fake clock,
K = 400 loop iterations, xrand_gen=10):
  plan_*   : tree.state, pID, per-node edge lengths, T, node_seq, plan states/efforts
  calls_*  : 64 random (x, u, xgoal) and the reference's own dynamics / lqr / erf /
             is_feasible outputs on them
  hot_*    : on that final tree, 32 fresh samples with the reference's _costs_to_go
             minima/argmins (pruning off/on) and _steer results
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

DEMOS = [('demo_boat_novice', 2, 0.5), ('demo_car', 5, 0), ('demo_boat_intermediate', 2, 0.5),
         ('demo_boat_advanced', 2, 0.9)]
K = 400


def main():
    for name, horizon, FPR in DEMOS:
        planner, ns = ref_shim.run_reference_demo(name, 0, K, dict(horizon=horizon, dt=0.1, FPR=FPR))
        tree = planner.tree
        n, m = ns['nstates'], ns['ncontrols']
        out = dict(K=K, seed=0, state=tree.state, pID=np.array(tree.pID), T=planner.T,
                   reached=planner.plan_reached_goal, node_seq=np.array(planner.node_seq),
                   x_seq=np.array(planner.x_seq), u_seq=np.array(planner.u_seq),
                   edge_len=np.array([len(e) for e in tree.x_seq]),
                   obs=np.array(ns['obs'], dtype=np.float64))
        # ---- callback vectors
        rng = np.random.RandomState(123)
        ss = np.array(ns['sample_space'], dtype=np.float64)
        lo = np.minimum(ss[:, 0], -1.0) - 1.0; hi = np.maximum(ss[:, 1], 1.0) + 1.0
        X = lo + (hi - lo) * rng.rand(64, n)
        XG = lo + (hi - lo) * rng.rand(64, n)
        U = 3000 * (rng.rand(64, m) - 0.5)
        # states drawn from the grown tree hit the feasibility boundary far more often
        X[:32] = tree.state[rng.randint(0, tree.size, 32)] + 0.3 * rng.randn(32, n)
        out['cb_x'] = X; out['cb_u'] = U; out['cb_xg'] = XG
        out['cb_dyn'] = np.array([ns['dynamics'](np.copy(x), np.copy(u), 0.1) for x, u in zip(X, U)])
        out['cb_K'] = np.array([ns['lqr'](x, u)[1] for x, u in zip(X, U)], dtype=np.float64)
        out['cb_S'] = np.array([ns['lqr'](x, u)[0] for x, u in zip(X, U)], dtype=np.float64)
        out['cb_erf'] = np.array([ns['erf'](np.copy(xg), np.copy(x)) for xg, x in zip(XG, X)])
        out['cb_feas'] = np.array([bool(ns['is_feasible'](x, u)) for x, u in zip(X, U)])
        # ---- hot-path vectors on the final tree
        Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(32, n)
        costs = np.array([planner._costs_to_go(q) for q in Q])
        out['hot_q'] = Q
        out['hot_argmin'] = np.argmin(costs, axis=1)
        out['hot_mincost'] = np.min(costs, axis=1)
        ign = np.unique(np.array(planner.node_seq)) if planner.plan_reached_goal else np.array([], dtype=np.int64)
        out['hot_ignores'] = ign
        pr = []
        for c in costs:
            order = np.argsort(c)
            pick = order[0]
            for ID in order:
                if ID not in ign:
                    pick = ID
                    break
            pr.append(pick)
        out['hot_pruned'] = np.array(pr)
        H = int(horizon / 0.1)
        sx = np.zeros((32, H, n)); su = np.zeros((32, H, m)); sl = np.zeros(32, dtype=np.int64)
        for j, q in enumerate(Q):
            xs, us = planner._steer(int(out['hot_argmin'][j]), q, force_arrive=False)
            sl[j] = len(xs)
            if len(xs):
                sx[j, :len(xs)] = np.array(xs); su[j, :len(xs)] = np.array(us)
        out['hot_steer_x'] = sx; out['hot_steer_u'] = su; out['hot_steer_len'] = sl
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'size', tree.size, 'T', planner.T, 'steer lens', sl.tolist()[:10], '->', os.path.getsize(path), 'bytes')


def adaptive():
    """Adaptive-horizon heuristic (planner.py:418-425): boat_novice with horizon=(0.1, 3) as the ROS
    behaviours use it (behaviors/params.py:31), two consecutive plans so the carried-over horizon_iters matters."""
    ref = ref_shim.import_reference()
    ns = ref_shim.load_demo_model('demo_boat_novice', 0)
    cons = ref.Constraints(nstates=6, ncontrols=3, goal_buffer=ns['goal_buffer'], is_feasible=ns['is_feasible'])
    clock = ref_shim.FrameClock()
    planner = ref.Planner(ns['dynamics'], ns['lqr'], cons, horizon=(0.1, 3), dt=0.1, FPR=0.5, error_tol=ns['error_tol'],
                          erf=ns['erf'], min_time=300, max_time=300, max_nodes=1E5, goal0=ns['goal'], sys_time=clock,
                          printing=False)
    out = {}
    for k in range(2):
        clock.t = -1
        planner.update_plan(ns['x0'], ns['sample_space'], goal_bias=ns['goal_bias'], xrand_gen=10)
        out['state%d' % k] = planner.tree.state
        out['pID%d' % k] = np.array(planner.tree.pID)
        out['edge_len%d' % k] = np.array([len(e) for e in planner.tree.x_seq])
        out['horizon_iters%d' % k] = planner.horizon_iters
        out['T%d' % k] = planner.T
    path = os.path.join(HERE, 'adaptive_boat_novice.npz')
    np.savez_compressed(path, **out)
    print('adaptive', [out['state%d' % k].shape for k in range(2)], [int(out['horizon_iters%d' % k]) for k in range(2)],
          'max edge', [int(out['edge_len%d' % k].max()) for k in range(2)])


if __name__ == '__main__':
    main()
    adaptive()
