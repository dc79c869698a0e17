"""Host-side plumbing for the two multi-GPU modes; torch.distributed is used only to move the
128-byte NCCL unique id and to barrier -- the gradient allreduce itself is issued by the engine on its stream.

  data parallel (config 4): contiguous row shards of one global batch, one sum-allreduce of the flat
      gradient (+ loss partial sums in its tail) per step, identical fused Adam on every rank;
  grid sharding (config 3): independent settings round-robin over ranks, no collective.
"""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def shard_rows(global_batch, rank, world):
    """(first_row, rows) of this rank's contiguous slice; the Philox streams are indexed by first_row + local
    row, so the union of the shards reproduces the 1-rank masks bit for bit."""
    if global_batch % world != 0:
        raise ValueError('global batch %d is not divisible by world size %d' % (global_batch, world))
    rows = global_batch // world
    return rank * rows, rows


def exchange_unique_id(make_id, rank, dist, device=None):
    """Rank 0 creates the id (make_id() -> 128 bytes); everyone receives it through a broadcast."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device or 'cpu')
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def attach_data_parallel(engine, global_batch, dist=None):
    """Creates the engine's NCCL communicator over the ranks of an initialised torch.distributed group and
    describes this rank's shard.  Returns (first_row, rows)."""
    import torch
    from .engine import Engine
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    first, rows = shard_rows(global_batch, rank, world)
    if world > 1:
        uid = exchange_unique_id(Engine.comm_unique_id, rank, dist, device=torch.device('cuda', torch.cuda.current_device()))
        engine.comm_init(uid, rank, world)
    engine.set_shard(global_batch, first)
    return first, rows
