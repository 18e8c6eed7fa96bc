"""Channel sharding across the GPUs of one box.

Channels/voices are independent (each is its own stateful_lambda in reference terms,
flowz/flowz.hpp:1181-1230), so rank r of G owns the contiguous range channel_range(C, G, r) with its
own plan, state and parameter slices, and the evaluation itself needs no communication.  The only
exchange is at the edges, when a block lives on one rank: ONE scatter of the input block and ONE
gather of the output block (NCCL grouped send/recv under torch.distributed; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def channel_range(channels: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of rank's channels: contiguous, sizes differ by at most one, first ranks larger."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(channels, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_sizes(channels: int, world_size: int) -> List[int]:
    return [channel_range(channels, world_size, r)[1] - channel_range(channels, world_size, r)[0]
            for r in range(world_size)]


def scatter_channels(block, channels: int, samples: int, root: int = 0, group=None, device=None):
    """Planar [channels, samples] block on `root` -> this rank's [own, samples] shard.
    `block` is only read on root (pass None elsewhere).  One grouped exchange."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = channel_range(channels, world, rank)
    if device is None:
        device = block.device if block is not None else torch.device("cpu")
    own = torch.empty((e - b, samples), dtype=torch.float32, device=device)
    if world == 1:
        own.copy_(block)
        return own
    ops = []
    if rank == root:
        for r in range(world):
            rb, re = channel_range(channels, world, r)
            if r == root:
                own.copy_(block[rb:re])
            elif re > rb:
                ops.append(dist.P2POp(dist.isend, block[rb:re].contiguous(), r, group))
    elif e > b:
        ops.append(dist.P2POp(dist.irecv, own, root, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return own


def gather_channels(own, channels: int, samples: int, root: int = 0, group=None, out: Optional[object] = None):
    """Inverse of scatter_channels: returns the [channels, samples] block on root, None elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        if out is None:
            return own.clone()
        out.copy_(own)
        return out
    ops = []
    if rank == root:
        if out is None:
            out = torch.empty((channels, samples), dtype=torch.float32, device=own.device)
        for r in range(world):
            rb, re = channel_range(channels, world, r)
            if r == root:
                out[rb:re].copy_(own)
            elif re > rb:
                ops.append(dist.P2POp(dist.irecv, out[rb:re], r, group))
    elif own.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, own.contiguous(), root, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out if rank == root else None
