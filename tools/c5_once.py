#!/usr/bin/env python
"""
ncu target for the FD-Jacobian + Riccati provider: one C5 plan (12-state double integrator, 5000 spheres) grown to
C5_NODES nodes with waves of C5_WAVE samples, so that steer_x_kernel<..FdRiccatiLqr..> and metric_coop_kernel run at a
working point that finishes in seconds.

    ncu --set full --clock-control none --import-source on -k regex:'steer_x|metric_coop' -s 30 -c 4 -o gpurun_out/ncu_riccati python tools/c5_once.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lqrrt_b200 import configs
    torch.cuda.set_device(0)
    prob = configs.BY_NAME[os.environ.get('C5_CONFIG', 'c5')](0)
    W = int(os.environ.get('C5_WAVE', '2048'))
    planner = prob.make_planner(dtype='f32', wave=W, max_nodes=int(os.environ.get('C5_NODES', '20000')), min_time=1e9, max_time=1e9, seed=1)
    planner.profile = True
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    prof = planner._engine.profile_read()
    w = max(1, planner.stats['waves'])
    print("nodes %d  samples %d  waves %d  device %.2f ms  per wave (us): %s" % (
        planner.tree.size, planner.stats['samples'], w, planner.stats['device_ms'], {k: round(v / w * 1e3, 1) for k, v in prof['ms'].items()}))


if __name__ == '__main__':
    main()
