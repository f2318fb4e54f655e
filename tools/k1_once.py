#!/usr/bin/env python
"""One launch each of the packed and the scalar nearest scan at the C4 working point (n = 6, W = 8192,
N = 1e5) -- the target of `ncu --set full -k regex:nearest_diag -c 2` (profiles/ncu_nearest_*_r1.txt)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from lqrrt_b200 import configs
from gpu_util import make_engine

torch.cuda.set_device(0)
prob = configs.c4_boat_advanced(0)
N, W = 100000, 8192
rng = np.random.RandomState(1)
ss = np.array(prob.sample_space)
nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(N, 6); nodes[:, 2] = 12 * (rng.rand(N) - 0.5)
Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(W, 6); Q[:, 2] = 12 * (rng.rand(W) - 0.5)
eng = make_engine(prob, 'f32', node_capacity=N + 8, wave_capacity=W)
eng.load_nodes(nodes)
q = torch.tensor(np.ascontiguousarray(Q.T), dtype=torch.float32, device='cuda')
idx = torch.empty(W, dtype=torch.int32, device='cuda'); cost = torch.empty(W, dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for kid in (1, 0):
    eng.set_option(1, kid)
    eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), cost.data_ptr(), st)
    torch.cuda.synchronize()
print("ok")
