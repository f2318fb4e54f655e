"""Multi-rank host logic on CPU: two processes over gloo exchange their per-wave record segments with the
same Exchange class the NCCL path uses, and agree on rank order (= sample order) and on rank 0's stop flag."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from lqrrt_b200 import waves
    seg = 256
    ex = waves.Exchange(dist.group.WORLD, seg, torch.device('cpu'))
    results = []
    stop_request = waves.STOP_NONE
    for wave in range(5):
        W = waves.wave_size(100 + 50 * wave, None, wave_max=96)
        first, count = waves.partition(W, rank, world)
        send = torch.zeros(seg, dtype=torch.uint8)
        hdr = send[:16].view(torch.int32)
        hdr[0] = count; hdr[1] = first; hdr[3] = stop_request      # header: count, (first), flags
        send[16:16 + count] = torch.arange(first, first + count, dtype=torch.uint8)
        recv = ex.gather(send)
        segs = recv.view(world, seg)
        order = torch.cat([segs[r, 16:16 + int(segs[r, :16].view(torch.int32)[0])] for r in range(world)])
        flags = int(segs[0, :16].view(torch.int32)[3])
        results.append((W, order.tolist(), flags))
        # rank 0 decides to stop after wave 2; everyone must see it in wave 3's header
        stop_request = waves.STOP_DONE if (rank == 0 and wave >= 2) else waves.STOP_NONE
        if flags != waves.STOP_NONE:
            break
    out[rank] = results
    dist.destroy_process_group()


def test_two_rank_exchange_order_and_stop_protocol():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a == b                                   # identical view on every rank
    assert len(a) == 4                              # waves 0..3: the flag raised after wave 2 lands in wave 3
    for W, order, flags in a:
        assert order == list(range(W))              # rank-order concatenation == sample order
    assert [f for _, _, f in a] == [0, 0, 0, 1]
