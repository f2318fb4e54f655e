"""
Host-side wave scheduling: how many samples the next launch expands, which
slice of them this GPU takes, and how the per-wave records of all GPUs are
exchanged.  Pure host logic (no CUDA calls) so it is unit-testable on CPU with
the gloo backend.

The reference expands one sample per loop iteration (lqrrt/planner.py:233-283);
a wave expands `W` of them against the tree as of wave start (SURVEY.md section 7
"what wave batching does to the planner").
"""
from __future__ import annotations


def wave_size(tree_size, fixed=None, wave_min=32, wave_max=8192, growth=1.0):
    """Global samples in the next wave.  `fixed` pins it (1 = the reference's
    sequential semantics).  Otherwise the wave tracks the tree size, so the tree
    at most roughly doubles per wave while it is small -- tree depth grows by at
    most one edge per wave, and a huge first wave would only fan out from the
    seed -- and saturates at wave_max.  `growth` scales it: W = growth * N; smaller values trade
    throughput for tree depth per sample."""
    if fixed is not None:
        return max(1, int(fixed))
    return int(min(wave_max, max(wave_min, growth * tree_size)))


def partition(wave, rank, world):
    """Contiguous slice [first, first+count) of a wave for `rank`: rank-order
    concatenation of the slices is sample order, which is what makes the
    G-GPU tree bit-identical to the 1-GPU tree for the same wave sizes."""
    base, extra = divmod(int(wave), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def local_capacity(wave_max, world):
    """Largest slice any rank can receive, rounded to the engine's granule."""
    per = -(-int(wave_max) // int(world))
    return (per + 3) & ~3


class Exchange(object):
    """All-gather of the fixed-size record segment every rank packs per wave
    (csrc/commit.cuh).  One NCCL collective per wave on the current stream;
    payload = accepted-node records, a few MB at most -- latency-, not
    bandwidth-bound, so one unsplit collective is the right granularity."""

    def __init__(self, group, segment_bytes, device):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.segment_bytes = int(segment_bytes)
        self.recv = torch.empty(self.world * self.segment_bytes, dtype=torch.uint8, device=device)

    def gather(self, send):
        """send: uint8 tensor of segment_bytes (a zero-copy view of the engine's
        record buffer).  Returns the (world * segment_bytes) tensor, rank-major."""
        if send.numel() != self.segment_bytes:
            raise ValueError("record segment has {} bytes, expected {}".format(send.numel(), self.segment_bytes))
        self.dist.all_gather_into_tensor(self.recv, send, group=self.group)
        return self.recv


class PeerExchange(object):
    """Fused exchange: every rank's receive buffer lives in peer-mapped symmetric memory; the engine's pack
    kernel stores records straight into all peers' buffers over NVLink and its commit kernel waits on
    arrival flags (csrc/commit.cuh).  No collective call per wave; this class only allocates and maps."""

    def __init__(self, group, segment_bytes, device, engine):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # four wave parities of world segments each, then the arrival flags (csrc/engine.cuh PEER_PAR)
        self.nbytes = 4 * self.world * int(segment_bytes) + 4 * self.world * 4
        self.nbytes = (self.nbytes + 255) & ~255
        self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                   # every rank's flags are zero before anyone stores
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        engine.set_peers(self.world, self.rank, self.ptrs, self.nbytes)


STOP_NONE, STOP_DONE, STOP_KILL = 0, 1, 2
