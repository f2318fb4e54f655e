#!/usr/bin/env python
"""
A/B measurements of the engine's issue-level options on one B200 (run under gpurun; writes JSON lines to
gpurun_out/ab_r1.jsonl as it goes, so a cut-off run still leaves what it finished):

  k1     the nearest scan alone, scalar vs packed FP32x2 kernel, bit-equality + CUDA-event time per launch
         at (C4: n=6, W=8192, N=1e5) and (C5: n=12, W=8192, N=1e6)
  plan   whole plans (bench.py's step) for {scalar, packed} x {prefetch off, on} x wave caps, nodes/s on the
         device clock and per-stage milliseconds

Every timed number is CUDA events on the launching stream after warm-up; nothing here runs under a profiler.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = os.path.join(ROOT, 'gpurun_out', 'ab_r1.jsonl')


def emit(rec):
    rec['t'] = round(time.time() - T0, 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, 'a') as f:
        f.write(json.dumps(rec) + '\n')
    print(json.dumps(rec), flush=True)


def k1_case(torch, cfg_name, N, W, reps=20):
    from lqrrt_b200 import configs
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
    res = {}
    for kname, kid in (('scalar', 0), ('packed', 1)):
        eng.set_option(1, kid)
        for _ in range(3):
            eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), cost.data_ptr(), st)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), cost.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[kname] = (ms, idx.cpu().numpy().copy(), cost.cpu().numpy().copy())
    same = bool(np.array_equal(res['scalar'][1], res['packed'][1]) and
                np.array_equal(res['scalar'][2].view(np.int32), res['packed'][2].view(np.int32)))
    pairs = float(W) * N
    emit({"what": "k1", "config": cfg_name, "n": n, "W": W, "N": N, "bit_identical": same,
          "ms_scalar": res['scalar'][0], "ms_packed": res['packed'][0],
          "pairs_per_s_scalar": pairs / res['scalar'][0] * 1e3, "pairs_per_s_packed": pairs / res['packed'][0] * 1e3,
          "note": "time includes the finalize kernel (~2 us)"})
    eng.close()


def plan_case(torch, cfg_name, nearest, prefetch, wave_max, growth, steps=5, warmup=2):
    from lqrrt_b200 import configs
    prob = configs.BY_NAME[cfg_name](0)
    planner = prob.make_planner(dtype='f32', wave=None, wave_max=wave_max, wave_growth=growth, max_nodes=prob.max_nodes,
                                min_time=1e9, max_time=1e9, nearest_kernel=nearest, prefetch=prefetch)
    planner.profile = True
    dev = []; host = []; nodes = []; waves = []; samples = []
    for s in range(warmup + steps):
        planner.seed = 9000 + s if s < warmup else s - warmup
        if s == warmup:
            planner._engine.profile_read(reset=True)
        t0 = time.perf_counter()
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
        h = time.perf_counter() - t0
        if s >= warmup:
            st = planner.stats
            dev.append(st['device_ms']); host.append(h * 1e3); nodes.append(st['nodes'] - 1); waves.append(st['waves'])
            samples.append(st['samples'])
    prof = planner._engine.profile_read()
    digest = int(np.frombuffer(np.array(planner.tree.state, dtype=np.float64).tobytes(), dtype=np.uint64).sum() & 0xFFFFFFFFFFFF)
    emit({"what": "plan", "config": cfg_name, "nearest": nearest, "prefetch": prefetch, "wave_max": wave_max, "growth": growth,
          "nodes_per_s": float(np.sum(nodes)) / (float(np.sum(dev)) / 1e3), "e2e_nodes_per_s": float(np.sum(nodes)) / (float(np.sum(host)) / 1e3),
          "ms_per_plan": float(np.mean(dev)), "waves": float(np.mean(waves)), "yield": float(np.sum(nodes)) / float(np.sum(samples)),
          "stage_ms_per_plan": {k: v / steps for k, v in prof['ms'].items()}, "pairs_per_s": prof['pairs'] / ((prof['ms']['nearest_old'] + prof['ms']['nearest_delta']) / 1e3),
          "last_tree_digest": digest, "depth": len(planner.node_seq)})
    planner._engine.close(); planner._engine = None


def main():
    import torch
    torch.cuda.set_device(0)
    emit({"what": "start", "gpu": torch.cuda.get_device_name(0)})
    which = sys.argv[1:] or ['k1', 'plan', 'wave', 'c5k']
    if 'k1' in which:
        k1_case(torch, 'c4', 100000, 8192)
        k1_case(torch, 'c4', 20000, 8192)
        k1_case(torch, 'c5k', 1000000, 8192, reps=5)
    if 'plan' in which:
        for nearest in ('scalar', 'packed'):
            for prefetch in (False, True):
                plan_case(torch, 'c4', nearest, prefetch, 8192, 1.0)
    if 'wave' in which:
        for wave_max, growth in ((16384, 1.0), (8192, 2.0), (16384, 2.0), (32768, 2.0), (8192, 4.0)):
            plan_case(torch, 'c4', 'packed', True, wave_max, growth)
    if 'c5k' in which:
        for nearest in ('scalar', 'packed'):
            plan_case(torch, 'c5k', nearest, True, 8192, 1.0, steps=1, warmup=1)
    emit({"what": "done"})


if __name__ == '__main__':
    T0 = time.time()
    main()
