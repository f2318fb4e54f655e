"""
Real multi-GPU path (NCCL over NVLink): one process per GPU, Planner(comm=True).  The G-rank tree must be
bit-identical on every rank and bit-identical to the single-GPU tree grown with the same global wave
schedule (parity level 4).  Needs >= 2 visible GPUs (gpurun --gpus 2).
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _grow(rank, world, port, out, dtype, exchange='peer', lag=2):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    from lqrrt_b200 import configs
    torch.cuda.set_device(rank)
    comm = None
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
        comm = True
    prob = configs.demo_boat_intermediate(0)
    planner = prob.make_planner(dtype=dtype, seed=11, wave_max=512, max_nodes=6000, min_time=1e9, max_time=1e9,
                                comm=comm, device=rank, exchange=exchange, lag=lag)
    # the wave schedule follows the tree size the host last saw, which must not depend on G
    planner.pipeline = True
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    t = planner.tree
    xs = np.array(planner.x_seq)
    out[(world, rank)] = dict(ok=ok, size=t.size, state=t.state.copy(), pID=list(t.pID), cost=t.cost.copy(), xs=xs,
                              waves=planner.stats['waves'], T=planner.T, reached=planner.plan_reached_goal,
                              node_seq=list(getattr(planner, 'node_seq', [])))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize('dtype,exchange,lag', [('f32', 'peer', 2), ('f64', 'peer', 2), ('f32', 'peer', 1), ('f32', 'nccl', 1)])
def test_two_gpu_tree_equals_one_gpu_tree(dtype, exchange, lag):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    # (the host-scheduled NCCL path has the lag-1 wave semantics; the 1-GPU tree to compare with is grown with the same lag)
    mp.spawn(_grow, args=(1, _free_port(), out, dtype, 'peer', lag), nprocs=1, join=True)
    mp.spawn(_grow, args=(world, _free_port(), out, dtype, exchange, lag), nprocs=world, join=True)
    one = out[(1, 0)]
    assert one['size'] == 6001 and one['ok'] is False          # stopped by max_nodes (planner.py:330-336)
    for r in range(world):
        g = out[(world, r)]
        assert g['size'] == one['size']
        assert one['waves'] <= g['waves'] <= one['waves'] + 3      # rank 0's stop flag lands a wave or two later
        assert g['pID'] == one['pID']
        assert np.array_equal(g['state'], one['state'])        # bitwise
        assert np.array_equal(g['cost'], one['cost'])
        assert g['reached'] == one['reached'] and g['T'] == one['T'] and g['node_seq'] == one['node_seq']
        assert np.array_equal(g['xs'], one['xs'])              # plan edges assembled from their owner ranks
