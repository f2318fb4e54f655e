#!/usr/bin/env python
"""
Sweep of the nearest-scan variants (include/lqrrt_b200.h LQ_NEAREST_*) x grid targets on one B200:
bit-equality against the scalar kernel and CUDA-event time per launch at the C4 and C5 working points,
then whole C4 plans for every variant.  JSON lines to gpurun_out/sweep_k1_r1.jsonl.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = os.path.join(ROOT, 'gpurun_out', 'sweep_k1_r1.jsonl')
VARIANTS = ['scalar', 'packed', 'packed_q8', 'two_level', 'prefix']   # round 1 also swept packed_r48 (48 registers) and packed_q2; never best, removed
# ('prefix' was added after round 1's last GPU call and has not been measured yet)
T0 = time.time()


def emit(rec):
    rec['t'] = round(time.time() - T0, 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, 'a') as f:
        f.write(json.dumps(rec) + '\n')
    print(json.dumps(rec), flush=True)


def k1(torch, cfg_name, N, W, reps, ctas_list):
    from lqrrt_b200 import configs
    from lqrrt_b200._native import NEAREST_KERNELS
    from gpu_util import make_engine
    prob = configs.BY_NAME[cfg_name](0)
    n = prob.nstates
    rng = np.random.RandomState(1)
    ss = np.array(prob.sample_space)
    nodes = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(N, n)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(W, n)
    if cfg_name == 'c4':
        nodes[:, 2] = 12 * (rng.rand(N) - 0.5); Q[:, 2] = 12 * (rng.rand(W) - 0.5)
    eng = make_engine(prob, 'f32', node_capacity=N + 8, wave_capacity=W)
    eng.load_nodes(nodes)
    q = torch.tensor(np.ascontiguousarray(Q.T), dtype=torch.float32, device='cuda')
    idx = torch.empty(W, dtype=torch.int32, device='cuda'); cost = torch.empty(W, dtype=torch.float32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    for name in VARIANTS:
        for ctas in ctas_list:
            eng.set_option(1, NEAREST_KERNELS[name]); eng.set_option(3, ctas)
            for _ in range(2):
                eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), cost.data_ptr(), st)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), cost.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            got = (idx.cpu().numpy().copy(), cost.cpu().numpy().view(np.int32).copy())
            if ref is None:
                ref = got
            same = bool(np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]))
            emit({"what": "k1", "config": cfg_name, "n": n, "W": W, "N": N, "kernel": name, "ctas_per_sm": ctas, "ms": ms,
                  "pairs_per_s": float(W) * N / ms * 1e3, "bit_identical_to_scalar": same})
    eng.close()


def plan(torch, name, ctas, steps=4, warmup=2, growth=2.0):
    from lqrrt_b200 import configs
    prob = configs.c4_boat_advanced(0)
    planner = prob.make_planner(dtype='f32', wave=None, wave_max=8192, wave_growth=growth, max_nodes=prob.max_nodes,
                                min_time=1e9, max_time=1e9, nearest_kernel=name, prefetch=True)
    planner.profile = True
    dev = []; nodes = []
    for s in range(warmup + steps):
        planner.seed = 9000 + s if s < warmup else s - warmup
        if s == 0:
            planner._ensure_engine().set_option(3, ctas)
        if s == warmup:
            planner._engine.profile_read(reset=True)
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
        if s >= warmup:
            dev.append(planner.stats['device_ms']); nodes.append(planner.stats['nodes'] - 1)
    prof = planner._engine.profile_read()
    emit({"what": "plan", "kernel": name, "ctas_per_sm": ctas, "growth": growth, "nodes_per_s": float(np.sum(nodes)) / (float(np.sum(dev)) / 1e3),
          "ms_per_plan": float(np.mean(dev)), "nearest_ms_per_plan": (prof['ms']['nearest_old'] + prof['ms']['nearest_delta']) / steps,
          "pairs_per_s": prof['pairs'] / ((prof['ms']['nearest_old'] + prof['ms']['nearest_delta']) / 1e3)})
    planner._engine.close(); planner._engine = None


def main():
    import torch
    torch.cuda.set_device(0)
    emit({"what": "start", "gpu": torch.cuda.get_device_name(0)})
    k1(torch, 'c4', 100000, 8192, 10, (8, 12, 16))
    k1(torch, 'c4', 30000, 8192, 10, (8, 16))
    k1(torch, 'c5k', 1000000, 8192, 3, (8, 16))
    for name in VARIANTS:
        for ctas in (8, 16):
            plan(torch, name, ctas)
    emit({"what": "done"})


if __name__ == '__main__':
    main()
