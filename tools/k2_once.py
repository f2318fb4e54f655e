#!/usr/bin/env python
"""
ncu target for the rollout kernel (K2): grows a C4 tree to ~40k nodes through the planner, then launches the
stand-alone rollout entry point (lq_steer: same rollout core and shared-memory table staging as the wave kernel)
for 8192 / 1024 (node, target) pairs drawn the way a wave draws them, for each lane configuration.

    ncu --set full --clock-control none --import-source on -k regex:steer_kernel -c 6 -o gpurun_out/k2 python tools/k2_once.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    from lqrrt_b200 import configs
    from gpu_util import nearest
    torch.cuda.set_device(0)
    prob = configs.c4_boat_advanced(0)
    planner = prob.make_planner(dtype='f32', wave_max=8192, wave_growth=2.0, max_nodes=40000, min_time=1e9, max_time=1e9, seed=1)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    eng = planner._engine
    rng = np.random.RandomState(0)
    ss = np.array(prob.sample_space)
    for W in (8192, 1024):
        Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(W, 6)
        Q[rng.rand(W) < 0.2, :2] = prob.goal[:2]
        ids, _ = nearest(eng, Q)
        q = torch.tensor(np.ascontiguousarray(Q.T), dtype=torch.float32, device='cuda')
        i = torch.tensor(ids, dtype=torch.int32, device='cuda')
        ln = torch.empty(W, dtype=torch.int32, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        for lanes in (8, 1):
            eng.set_tuning(lanes, 0)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            eng.steer(i.data_ptr(), q.data_ptr(), W, None, None, ln.data_ptr(), st)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                eng.steer(i.data_ptr(), q.data_ptr(), W, None, None, ln.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            L = ln.cpu().numpy()
            print("W %5d lanes %2d: %.1f us per launch, edge length mean %.1f max %d, accepted %.2f" % (
                W, lanes, e0.elapsed_time(e1) / 5 * 1e3, L.mean(), L.max(), (L > 0).mean()), flush=True)


if __name__ == '__main__':
    main()
