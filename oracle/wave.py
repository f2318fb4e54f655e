"""
ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/models.py header).

Wave-batched CPU oracle: the semantics the CUDA engine implements when more
than one sample is expanded per launch (parity level 3).  RRT is
sequential in the reference (planner.py:233-290: iteration k searches a tree
that already holds iteration k-1's node), so a wave is a *different planner*
for W > 1 and has no reference to pin against -- what is pinned is that W = 1
reproduces the sequential oracle (tests/test_wave_oracle.py), which is itself
pinned bit-for-bit to the reference.

Wave semantics (identical in csrc/engine.cu):
  1. every sample of the wave is drawn (Philox, oracle/philox.py, or supplied)
  2. every sample's nearest node is searched in the tree *as of wave start*
     (planner.py:239-247,340-350 per sample; lowest index wins ties)
  3. every sample is steered from that node (planner.py:354-438)
  4. samples whose steer kept >= 1 state are appended in sample order
     (planner.py:253-257, tree.py:77-96)
  5. goal tests / ignore marks / best plan are applied in node order
     (planner.py:260-281), so they affect the *next* wave's searches.

Pipelined waves (`lag = 2`, the device-scheduled driver's default): wave k searches the tree -- nodes AND
ignore marks -- as committed by wave k-2 instead of k-1, so that the rollouts of two consecutive waves are
independent of each other and can be in flight together; everything else is unchanged (nodes are still
appended wave after wave in sample order).  `expand_wave(..., lag=2)` keeps the snapshots.
"""
import numpy as np

from . import philox
from .models import make_callbacks, make_is_feasible_batch


class WaveOracle(object):
    def __init__(self, system, goal, goal_buffer, horizon, dt, FPR, error_tol, node_capacity=1 << 30):
        self.sys = system
        self.n = system.nstates; self.m = system.ncontrols
        self.dynamics, self.lqr, self.erf, self.is_feasible = make_callbacks(system, dt, all_obstacles_at_once=True)
        self.feas_batch = make_is_feasible_batch(system.is_feasible)
        self.angle_dims = tuple(system.erf.angle_dims)
        self.dt = dt; self.FPR = FPR
        if hasattr(horizon, '__contains__'):          # adaptive-horizon heuristic, planner.py:538-547
            self.hspan = np.divide(horizon, dt).astype(np.int64)
            self.hfactor = 2
            self.horizon_iters = 1
            self.H = int(self.hspan[1])
        else:
            self.H = int(horizon / dt)
            self.hspan = (self.H, self.H); self.hfactor = 0; self.horizon_iters = self.H
        self.error_tol = np.abs(error_tol).astype(np.float64) * np.ones(self.n)
        self.goal = np.array(goal, dtype=np.float64)
        gb = np.abs(goal_buffer).astype(np.float64)
        self.goal_lo = self.goal - gb; self.goal_hi = self.goal + gb
        self.capacity = node_capacity

    # -- tree ---------------------------------------------------------------
    def reset(self, x0):
        x0 = np.array(x0, dtype=np.float64)
        self.state = x0.reshape(1, self.n).copy()
        self.parent = [-1]
        self.cost = [0]                      # cumulative recorded steps; T = (1 + cost) dt
        self.ignore = [False]
        self.K = [np.array(self.lqr(x0, np.zeros(self.m))[1], dtype=np.float64)]
        self.x_seq = [[x0]]; self.u_seq = [[np.zeros(self.m)]]
        self.best_node = -1; self.best_cost = None
        self.goal_nodes = []
        self.__dict__.pop('_snapshots', None)

    def load(self, state, parent, cost=None, K=None):
        """Inject a frozen tree (level-1 tests)."""
        self.state = np.array(state, dtype=np.float64)
        N = len(self.state)
        self.parent = list(parent)
        self.cost = [0] * N if cost is None else list(cost)
        self.ignore = [False] * N
        zero = np.zeros(self.m)
        self.K = [np.array(self.lqr(x, zero)[1], dtype=np.float64) for x in self.state] if K is None else list(K)
        self.x_seq = [[x] for x in self.state]; self.u_seq = [[zero] for _ in range(N)]
        self.best_node = -1; self.best_cost = None
        self.goal_nodes = []

    @property
    def size(self):
        return len(self.state)

    # -- 1. sampling (planner.py:194-211 on the Philox stream) ---------------
    def sample_wave(self, seed, wave, first, count, sample_space, goal_bias, tries):
        ss = np.array(sample_space, dtype=np.float64)
        centers = np.mean(ss, axis=1); spans = np.diff(ss).flatten()
        bias = np.array(goal_bias, dtype=np.float64) * np.ones(self.n)
        idx = np.arange(first, first + count, dtype=np.uint64)
        X = np.zeros((count, self.n)); done = np.zeros(count, dtype=bool)
        used = np.zeros(count, dtype=np.int64)
        zero_u = np.zeros((count, self.m))
        for t in range(tries):
            u = philox.uniforms(seed, wave, idx, t, self.n + 1)
            cand = centers + spans * (u[:, :self.n] - 0.5)
            pick = bias[None, :] > u[:, self.n:self.n + 1]
            cand = np.where(pick, self.goal[None, :], cand)
            ok = self.feas_batch(cand, zero_u)
            X[~done] = cand[~done]
            used[~done] = t + 1
            done |= ok
            if done.all():
                break
        return X, used

    # -- 2. nearest (planner.py:340-350 + :239-247) ---------------------------
    def diffs(self, x):
        """erf_v(x, tree.state) of planner.py:585-588 for all nodes at once."""
        d = x[None, :] - self.state
        for a in self.angle_dims:
            v = self.state[:, a]; g = x[a]
            c = np.cos(g); s = np.sin(g); cv = np.cos(v); sv = np.sin(v)
            d[:, a] = -np.arctan2(sv * c - cv * s, cv * c + sv * s)
        return d

    def costs(self, x):
        S = np.array(self.lqr(x, np.zeros(self.m))[0], dtype=np.float64)
        d = self.diffs(x)
        return np.sum(d.dot(S) * d, axis=1)

    def nearest(self, X, use_ignore=True, visible=None):
        """`visible` = (size, ignore list) snapshot the search is restricted to (pipelined waves); None = the tree as it is."""
        ids = np.zeros(len(X), dtype=np.int64); best = np.zeros(len(X))
        nvis = self.size if visible is None else visible[0]
        ign = np.array(self.ignore if visible is None else visible[1], dtype=bool)[:nvis]
        for j, x in enumerate(X):
            c = self.costs(x)[:nvis]
            if use_ignore and not ign.all():
                c2 = np.where(ign, np.inf, c)
                ids[j] = int(np.argmin(c2))
            else:
                ids[j] = int(np.argmin(c))
            best[j] = c[ids[j]]
        return ids, best

    # -- 3. steer (planner.py:354-438) ----------------------------------------
    def steer(self, ID, xtar, votes=None):
        """Every rollout of a wave starts from the planner-level horizon as of wave start; with the adaptive
        heuristic (planner.py:418-425) its doublings / halving are returned through `votes` as
        (n_doublings, halved) and applied in sample order by commit."""
        K = np.copy(self.K[ID]); x = np.copy(self.state[ID])
        xs = []; us = []; i = 0
        hor = self.horizon_iters
        last_emag = np.inf
        doubles = 0; halved = False
        while True:
            e = self.erf(np.copy(xtar), np.copy(x))
            u = K.dot(e)
            x = self.dynamics(np.copy(x), np.copy(u), self.dt)
            if not self.is_feasible(x, u):
                keep = int(self.FPR * len(xs))
                xs = xs[:keep]; us = us[:keep]
                break
            i += 1
            emag = np.abs(e)
            if self.hfactor:
                if np.all(emag >= last_emag):
                    xs = []; us = []; halved = True
                    break
                if i == hor:
                    hor = int(np.clip(self.hfactor * hor, self.hspan[0], self.hspan[1])); doubles += 1
                last_emag = emag
            if i > hor or np.all(emag <= self.error_tol):
                break
            xs.append(x); us.append(u)
            K = np.array(self.lqr(x, u)[1], dtype=np.float64)
        if votes is not None:
            votes.append((doubles, halved))
        return xs, us

    def apply_votes(self, votes):
        for doubles, halved in votes:
            for _ in range(doubles):
                self.horizon_iters = int(np.clip(self.hfactor * self.horizon_iters, self.hspan[0], self.hspan[1]))
            if halved:
                self.horizon_iters = int(np.clip(self.horizon_iters / self.hfactor, self.hspan[0], self.hspan[1]))

    def in_goal(self, x):
        return bool(np.all((self.goal_lo < x) & (x < self.goal_hi)))

    # -- 4+5. commit ----------------------------------------------------------
    def commit(self, ids, edges):
        added = 0; first_new = self.size
        rows = []
        for ID, (xs, us) in zip(ids, edges):
            if len(xs) == 0 or self.size + added >= self.capacity:
                continue
            xnew = np.copy(xs[-1])
            rows.append(xnew)
            self.parent.append(int(ID))
            self.cost.append(self.cost[int(ID)] + len(xs))
            self.ignore.append(False)
            self.K.append(np.array(self.lqr(xnew, np.copy(us[-1]))[1], dtype=np.float64))
            self.x_seq.append(xs); self.u_seq.append(us)
            added += 1
        if rows:
            self.state = np.vstack((self.state, np.array(rows)))
        for node in range(first_new, first_new + added):
            if self.in_goal(self.state[node]):
                self.goal_nodes.append(node)
                k = node
                while k != -1:
                    self.ignore[k] = True
                    k = self.parent[k]
                if self.best_cost is None or self.cost[node] < self.best_cost:
                    self.best_cost = self.cost[node]; self.best_node = node
        return added

    def expand_wave(self, X, use_ignore=True, lag=1):
        visible = None
        if lag > 1:
            # snapshots after every commit; wave k sees the state after commit k - lag (the initial tree before that)
            hist = self.__dict__.setdefault('_snapshots', [(self.size, list(self.ignore))])
            visible = hist[max(0, len(hist) - lag)]
        ids, _ = self.nearest(X, use_ignore, visible)
        votes = []
        edges = [self.steer(int(i), x, votes) for i, x in zip(ids, X)]
        added = self.commit(ids, edges)
        if self.hfactor:
            self.apply_votes(votes)
        if lag > 1:
            self._snapshots.append((self.size, list(self.ignore)))
        return added, ids, edges

    def climb(self, ID):
        path = []
        while ID != -1:
            path.append(ID); ID = self.parent[ID]
        return path[::-1]

    @property
    def T(self):
        return np.inf if self.best_cost is None else (1 + self.best_cost) * self.dt
