#!/usr/bin/env python
"""
Round-2 measurement (needs a GPU): where does a rollout step's latency go?

Runs the stand-alone rollout entry point (lq_steer: planner.py:354-438 for W (node, target) pairs) on the C4
boat with features switched off one at a time -- obstacles, the heading "magic rudder" (atan2 + sincos + atan2,
demo_boat_advanced.py:101-108), the car-like clamps, the thruster mix -- for every lane configuration, and
reports microseconds per launch and per rollout step of the longest rollout.  Rollouts are set up to run the
whole horizon (far targets in free space), so kernel time / steps = per-step latency of the dependent chain.
JSON lines to gpurun_out/steer_probe.jsonl.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = os.path.join(ROOT, 'gpurun_out', 'steer_probe.jsonl')


def variants():
    from lqrrt_b200 import configs, models
    base = configs.c4_boat_advanced(0)
    p = models.boat_advanced_params()
    vps = models.vertex_cloud(models.BOAT_LENGTH, models.BOAT_WIDTH, 0.25, 1)
    lo = np.array([-np.inf, -np.inf, -np.inf, -0.65, -0.4, -0.2]); hi = np.array([np.inf, np.inf, np.inf, 1.1, 0.4, 0.2])

    def make(name, rudder=4000, carlike=True, clip='thruster', obstacles=True, box=True):
        kw = dict(invM=[1 / 500] * 3, D_pos=p['D_pos'], D_neg=p['D_neg'], rudder=rudder, carlike=carlike,
                  velmax_pos0=p['velmax_pos'][0], velmax_neg0=p['velmax_neg'][0])
        if clip == 'thruster':
            dyn = models.BoatDynamics(B=p['B'], invB=p['invB'], thrust_max=p['thrust_max'], clip='thruster', **kw)
        else:
            dyn = models.BoatDynamics(u_max=[600.0, 600.0, 1800.0], clip='wrench', **kw)
        prims = ([models.StateBox(lo, hi)] if box else []) + ([models.CirclesCloud(base.obs, vps)] if obstacles else [])
        system = models.System(dyn, models.PDRotatedLqr(np.diag([120, 20, 0]), np.diag([120, 20, 0])), models.Erf([2]),
                               models.Feasibility(prims), name=name)
        return configs.Problem(name, system, base.x0, base.goal, base.goal_buffer, base.error_tol, base.sample_space,
                               base.goal_bias, base.horizon, base.dt, base.FPR)
    yield make('c4 as benchmarked')
    yield make('no obstacle table', obstacles=False)
    yield make('no obstacle table, no velocity box', obstacles=False, box=False)
    yield make('no rudder term', rudder=0.0)
    yield make('no rudder, not car-like', rudder=0.0, carlike=False)
    yield make('wrench clip instead of thruster mix', clip='wrench')
    yield make('bare: no rudder, not car-like, wrench clip, no constraints', rudder=0.0, carlike=False, clip='wrench',
               obstacles=False, box=False)


def main():
    import torch
    from gpu_util import make_engine
    torch.cuda.set_device(0)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    W = 8192
    rng = np.random.RandomState(0)
    for prob in variants():
        eng = make_engine(prob, 'f32', node_capacity=4096, wave_capacity=W)
        # nodes: boats cruising at 1 m/s in obstacle-free water far outside the arena; targets 30 m ahead
        nodes = np.zeros((1024, 6)); nodes[:, 0] = -500 - 100 * rng.rand(1024); nodes[:, 1] = -500 - 100 * rng.rand(1024)
        nodes[:, 2] = 0.3 * rng.randn(1024); nodes[:, 3] = 1.0
        eng.load_nodes(nodes)
        ids = rng.randint(0, 1024, W)
        targets = nodes[ids].copy()
        targets[:, 0] += 30 * np.cos(nodes[ids, 2]); targets[:, 1] += 30 * np.sin(nodes[ids, 2])
        q = torch.tensor(np.ascontiguousarray(targets.T), dtype=torch.float32, device='cuda')
        i = torch.tensor(ids, dtype=torch.int32, device='cuda')
        ln = torch.empty(W, dtype=torch.int32, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        for lanes in (1, 8, 32):
            eng.set_tuning(lanes, 0)
            for _ in range(3):
                eng.steer(i.data_ptr(), q.data_ptr(), W, None, None, ln.data_ptr(), st)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                eng.steer(i.data_ptr(), q.data_ptr(), W, None, None, ln.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            L = ln.cpu().numpy()
            rec = {"model": prob.name, "lanes": lanes, "W": W, "us_per_launch": us, "steps_max": int(L.max()) + 1,
                   "steps_mean": float(L.mean()) + 1, "us_per_step_of_longest": us / (int(L.max()) + 1)}
            with open(OUT, 'a') as f:
                f.write(json.dumps(rec) + '\n')
            print(json.dumps(rec), flush=True)
        eng.close()


if __name__ == '__main__':
    main()
