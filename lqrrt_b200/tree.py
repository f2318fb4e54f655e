"""
Tree: host-side view of the node store that lives in HBM (API of reference
lqrrt/tree.py: state / pID / lqr / x_seq / u_seq / size, add_node, climb,
trajectory, visualize).

The reference keeps an (N x n) float64 ndarray grown by np.vstack plus Python
lists (tree.py:62-70, 87-93).  Here the arrays belong to the engine
(structure-of-arrays on the device, csrc/commit.cuh TreeView) and this class
downloads on demand: node rows incrementally, edges per request (with several
GPUs an edge is held by the rank that rolled it out, so reading edges is a
collective there: every rank asks for the same nodes).  Reading
`tree.state[ID]`, `tree.x_seq[ID]`, `tree.pID` etc. therefore works as in the
reference's demos (demo_car.py:288-295).
"""
import numpy as np


class _LazySeq(object):
    """List-like whose items are fetched from the device when indexed."""

    def __init__(self, tree, fetch):
        self._tree = tree
        self._fetch = fetch

    def __len__(self):
        return self._tree.size

    def _norm(self, ID):
        n = len(self)
        if ID < 0:
            ID += n
        if not 0 <= ID < n:
            raise IndexError("list index out of range")
        return ID

    def __getitem__(self, ID):
        if isinstance(ID, slice):
            ids = list(range(*ID.indices(len(self))))
            return self._fetch(ids)
        return self._fetch([self._norm(int(ID))])[0]

    def __iter__(self):
        n = len(self)
        for lo in range(0, n, 4096):
            for item in self._fetch(list(range(lo, min(n, lo + 4096)))):
                yield item


class Tree(object):
    """Bound to an engine by the Planner; ID == insertion order, seed is ID 0
    with pID -1 (tree.py:17-20)."""

    def __init__(self, engine, seed_state, system, comm=None, rank=0, world=1):
        self._eng = engine
        self._sys = system
        self._comm = comm; self._rank = rank; self._world = world
        self.nstates = engine.n
        self.ncontrols = engine.m
        self._seed = np.array(seed_state, dtype=np.float64)
        self._n_cached = 0
        self._state = np.zeros((0, self.nstates))
        self._parent = np.zeros(0, dtype=np.int32)
        self._cost = np.zeros(0, dtype=np.int32)
        self._elen = np.zeros(0, dtype=np.int32)
        self._size = 1
        self._long_edges = {}       # node -> (x_seq, u_seq) for edges kept on the host (force-arrive)
        self.x_seq = _LazySeq(self, lambda ids: self._edges(ids)[0])
        self.u_seq = _LazySeq(self, lambda ids: self._edges(ids)[1])
        self.lqr = _LazySeq(self, self._lqr)

    # -- size / sync ---------------------------------------------------------
    @property
    def size(self):
        return self._size

    def _set_size(self, size):
        self._size = int(size)

    def _sync(self):
        if self._n_cached < self._size:
            d = self._eng.download_nodes(self._n_cached, self._size - self._n_cached)
            self._state = np.vstack((self._state, d['state']))
            self._parent = np.concatenate((self._parent, d['parent']))
            self._cost = np.concatenate((self._cost, d['cost']))
            self._elen = np.concatenate((self._elen, d['edge_len']))
            self._n_cached = self._size

    @property
    def state(self):
        self._sync()
        return self._state

    @property
    def pID(self):
        self._sync()
        return self._parent.tolist()

    @property
    def cost(self):
        """Cumulative recorded steps from the seed; plan duration = (1 + cost) * dt."""
        self._sync()
        return self._cost

    # -- per-node features ------------------------------------------------------
    def _edges(self, ids):
        xs = [None] * len(ids); us = [None] * len(ids)
        need = []
        for k, ID in enumerate(ids):
            if ID == 0:                          # tree.py:69-70: seed edge is [seed], [zeros]
                xs[k] = [self._seed]; us[k] = [np.zeros(self.ncontrols)]
            elif ID in self._long_edges:
                xs[k], us[k] = self._long_edges[ID]
            else:
                need.append(k)
        if need:
            hx, hu, hl = self._eng.download_edges([ids[k] for k in need])
            if self._world > 1:
                # an edge lives on the GPU that rolled it out (csrc/commit.cuh); the other ranks answer
                # zeros, so the sum over ranks is the edge.  COLLECTIVE: every rank must ask for the same
                # ids, which holds for the plan read-out in update_plan and for code that all ranks run alike.
                import torch
                import torch.distributed as dist
                dev = torch.device('cuda', self._eng.device) if dist.get_backend(self._comm) == 'nccl' else torch.device('cpu')
                buf = torch.from_numpy(np.concatenate((hx.reshape(len(need), -1), hu.reshape(len(need), -1)), axis=1)).to(dev)
                dist.all_reduce(buf, group=self._comm)
                buf = buf.cpu().numpy()
                nx = self._eng.H * self.nstates
                hx = buf[:, :nx].reshape(hx.shape); hu = buf[:, nx:].reshape(hu.shape)
            for j, k in enumerate(need):
                L = int(hl[j])
                xs[k] = [hx[j, i].copy() for i in range(L)]
                us[k] = [hu[j, i].copy() for i in range(L)]
        return xs, us

    def _lqr(self, ids):
        self._sync()
        S, K = self._eng.eval_lqr(self._state[ids], np.zeros((len(ids), self.ncontrols)))
        Kn = np.zeros((len(ids), self.ncontrols, self.nstates))
        for k, ID in enumerate(ids):             # the stored gain (planner.py:257), the metric at the node
            Kn[k] = self._eng.download_nodes(ID, 1, want_K=True)['K'][0]
        return [(S[k], Kn[k]) for k in range(len(ids))]

    # -- reference methods ------------------------------------------------------
    def add_node(self, pID, state, lqr, x_seq, u_seq):
        """tree.py:77-96.  `lqr` is accepted for signature compatibility; the engine stores
        K = lqr(state, 0) itself."""
        if pID >= self.size or pID < 0:
            raise ValueError("The given parent ID, {}, doesn't exist.".format(pID))
        x_seq = [np.array(x, dtype=np.float64) for x in x_seq]
        u_seq = [np.array(u, dtype=np.float64) for u in u_seq]
        self._eng.append_node(pID, state, np.array(x_seq).reshape(-1, self.nstates),
                              np.array(u_seq).reshape(-1, self.ncontrols))
        if len(x_seq) > self._eng.H:
            self._long_edges[self._size] = (x_seq, u_seq)
        self._size += 1

    def nearest(self, x, use_ignore=False):
        """Nearest node(s) to the state(s) `x` in cost-to-go, the reference's `_costs_to_go` + selection
        (planner.py:239-247, 340-350; BASELINE.json calls it Tree.nearest): S = lqr(x, 0).S at the query,
        cost_i = erf(x, v_i)' S erf(x, v_i) over every node, lowest cost wins, lowest ID on exact ties.  With
        `use_ignore` nodes on finished plans are skipped unless every node is (planner.py:240-245).  Runs the
        engine's scan kernel (lq_nearest).  x: (n,) -> (ID, cost);  (W, n) -> (IDs, costs)."""
        import torch
        q = np.asarray(x, dtype=np.float64)
        single = q.ndim == 1
        q = q.reshape(-1, self.nstates)
        if q.shape[0] < 1:
            raise ValueError("Expected at least one query state.")
        dev = torch.device('cuda', self._eng.device)
        tdt = torch.float32 if self._eng.dtype == 'f32' else torch.float64
        ids = np.zeros(len(q), dtype=np.int64); costs = np.zeros(len(q))
        cap = self._eng.wave_capacity
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            for lo in range(0, len(q), cap):
                part = q[lo:lo + cap]
                dq = torch.tensor(np.ascontiguousarray(part.T), dtype=tdt, device=dev)       # dim-major, as the scan reads it
                di = torch.empty(len(part), dtype=torch.int32, device=dev)
                dc = torch.empty(len(part), dtype=tdt, device=dev)
                self._eng.nearest(dq.data_ptr(), len(part), bool(use_ignore), di.data_ptr(), dc.data_ptr(), st)
                ids[lo:lo + len(part)] = di.cpu().numpy()
                costs[lo:lo + len(part)] = dc.double().cpu().numpy()
        if single:
            return int(ids[0]), float(costs[0])
        return ids, costs

    def climb(self, ID):
        """Node IDs from the seed to ID (tree.py:100-117)."""
        if ID >= self.size or ID < 0:
            raise ValueError("The given ID, {}, doesn't exist.".format(ID))
        if self._n_cached < self._size:
            return self._eng.climb(ID)           # parents are still on the device: walk them there
        path = []
        while ID != -1:
            path.append(int(ID))
            ID = self._parent[ID]
        path.reverse()
        return path

    def plan(self, ID=None, xguide=None, colmask=None):
        """(node_seq, x_seq, u_seq) of the plan that ends in node ID -- climb + trajectory (tree.py:100-132) -- or, with
        ID None, in the node closest to `xguide` (planner.py:313-318), read out by one engine call."""
        ids, hx, hu, el = self._eng.plan_readout(ID, xguide, colmask)
        if self._world > 1:                     # an edge lives on the GPU that rolled it out: sum the per-rank answers
            import torch
            import torch.distributed as dist
            dev = torch.device('cuda', self._eng.device) if dist.get_backend(self._comm) == 'nccl' else torch.device('cpu')
            buf = torch.from_numpy(np.concatenate((hx.reshape(len(ids), -1), hu.reshape(len(ids), -1)), axis=1)).to(dev)
            dist.all_reduce(buf, group=self._comm)
            buf = buf.cpu().numpy()
            nx = self._eng.H * self.nstates
            hx = buf[:, :nx].reshape(hx.shape); hu = buf[:, nx:].reshape(hu.shape)
        x_full = []; u_full = []
        for k, node in enumerate(ids):
            if node == 0:                        # tree.py:69-70: the seed's edge is [seed], [zeros]
                x_full.append(self._seed); u_full.append(np.zeros(self.ncontrols))
            elif node in self._long_edges:
                x_full.extend(self._long_edges[node][0]); u_full.extend(self._long_edges[node][1])
            else:
                L = int(el[k])
                x_full.extend(hx[k, :L].copy()); u_full.extend(hu[k, :L].copy())      # rows of one copy per edge
        return ids, x_full, u_full

    def trajectory(self, IDs):
        """Concatenated edges of the listed nodes (tree.py:121-132)."""
        xs, us = self._edges(list(IDs))
        x_full = []; u_full = []
        for a, b in zip(xs, us):
            x_full.extend(a); u_full.extend(b)
        return (x_full, u_full)

    def visualize(self, dx, dy, node_seq=None):
        """(dx, dy) cross-section plot (tree.py:136-171); needs matplotlib."""
        print("\n...now plotting...")
        from matplotlib import pyplot as plt
        fig = plt.figure(); fig.suptitle('Tree'); plt.axis('equal')
        ax = fig.add_subplot(1, 1, 1)
        ax.set_xlabel('- State {} +'.format(dx)); ax.set_ylabel('- State {} +'.format(dy)); ax.grid(True)
        on_plan = set(node_seq or [])
        if self.size > 1:
            for ID, edge in enumerate(self.x_seq):
                e = np.array(edge)
                if ID in on_plan:
                    ax.plot(e[:, dx], e[:, dy], color='r', zorder=2)
                else:
                    ax.plot(e[:, dx], e[:, dy], color='0.75', zorder=1)
        ax.scatter(self.state[0, dx], self.state[0, dy], color='b', s=48)
        if node_seq:
            ax.scatter(self.state[node_seq[-1], dx], self.state[node_seq[-1], dy], color='r', s=48)
        print("Done! Close window to continue.\n")
        plt.show()
