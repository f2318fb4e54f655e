"""Helpers shared by the -m gpu tests: engine construction from a Problem, torch <-> engine plumbing."""
import numpy as np


def make_engine(prob, dtype='f64', node_capacity=4096, wave_capacity=1024, **kw):
    from lqrrt_b200._native import Engine
    H = int(prob.horizon / prob.dt)
    eng = Engine(prob.system, H, prob.dt, prob.FPR, prob.error_tol, node_capacity=node_capacity,
                 wave_capacity=wave_capacity, dtype=dtype, device=0, **kw)
    eng.set_goal(prob.goal, prob.goal_buffer)
    return eng


def torch_dtype(eng):
    import torch
    return torch.float32 if eng.dtype == 'f32' else torch.float64


def nearest(eng, Q, use_ignore=False):
    """Q: (W, n) float64 host -> (idx, cost) numpy, through the device-pointer entry point."""
    import torch
    W = len(Q)
    q = torch.tensor(np.ascontiguousarray(Q.T), dtype=torch_dtype(eng), device='cuda')       # n x W dim-major
    idx = torch.empty(W, dtype=torch.int32, device='cuda')
    cost = torch.empty(W, dtype=torch_dtype(eng), device='cuda')
    eng.nearest(q.data_ptr(), W, use_ignore, idx.data_ptr(), cost.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return idx.cpu().numpy().astype(np.int64), cost.cpu().numpy().astype(np.float64)


def steer(eng, ids, Q):
    import torch
    W = len(Q)
    q = torch.tensor(np.ascontiguousarray(Q.T), dtype=torch_dtype(eng), device='cuda')
    i = torch.tensor(np.asarray(ids), dtype=torch.int32, device='cuda')
    xs = torch.empty((W, eng.H, eng.n), dtype=torch_dtype(eng), device='cuda')
    us = torch.empty((W, eng.H, eng.m), dtype=torch_dtype(eng), device='cuda')
    ln = torch.empty(W, dtype=torch.int32, device='cuda')
    eng.steer(i.data_ptr(), q.data_ptr(), W, xs.data_ptr(), us.data_ptr(), ln.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return xs.cpu().numpy().astype(np.float64), us.cpu().numpy().astype(np.float64), ln.cpu().numpy().astype(np.int64)


def tree_arrays(eng):
    n = eng.tree_size()
    return eng.download_nodes(0, n, want_K=True)
