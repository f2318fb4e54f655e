#!/usr/bin/env python
"""
ncu target for the wave pipeline's chain kernels: one C4 plan with FIXED waves of 1024 samples (what one GPU of eight
expands per wave) grown to 60k nodes, so that steer_x / commit_x / the delta scan run at the multi-GPU working point.

    ncu --set full --clock-control none --import-source on -k regex:'steer_x|commit_x' -s 60 -c 4 -o gpurun_out/chain python tools/pipe_once.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lqrrt_b200 import configs
    torch.cuda.set_device(0)
    prob = configs.c4_boat_advanced(0)
    W = int(os.environ.get('PIPE_WAVE', '1024'))
    planner = prob.make_planner(dtype='f32', wave=W, max_nodes=int(os.environ.get('PIPE_NODES', '60000')), min_time=1e9, max_time=1e9, seed=1)
    planner.profile = True
    for _ in range(int(os.environ.get('PIPE_PLANS', '1'))):
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    prof = planner._engine.profile_read()
    w = max(1, planner.stats['waves'])
    print("waves %d  device %.2f ms  per wave (us): %s  host %s" % (
        w, planner.stats['device_ms'], {k: round(v / w * 1e3 / int(os.environ.get('PIPE_PLANS', '1')), 1) for k, v in prof['ms'].items()},
        planner.stats.get('host_pipeline_ms')))


if __name__ == '__main__':
    main()
