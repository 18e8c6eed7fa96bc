"""Channel sharding across the GPUs of one box.

Channels/voices are independent (each is its own stateful_lambda in reference terms,
flowz/flowz.hpp:1181-1230), so rank r of G owns the contiguous range channel_range(C, G, r) with its
own plan, state and parameter slices, and the evaluation itself needs no communication.  The only
exchange is at the edges, when a block lives on one rank: ONE scatter of the input block and ONE
gather of the output block (NCCL grouped send/recv under torch.distributed; gloo in the CPU tests).

On a box with NVLink / NVSwitch the edge step can also be FUSED into the kernels: share_from_root() maps the
root's input and output blocks into every rank's address space (CUDA IPC, peer memory), and each rank's
zg_process() then reads its channel rows from, and writes its results straight into, the root's HBM -- the TMA
tensor maps of the streaming kernels simply point at peer memory, tile by tile over NVLink, overlapped with the
arithmetic; no staging copy, no separate collective (process_on_root_block()).
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


class PeerBlock:
    """A planar fp32 [channels, samples] block in the ROOT's HBM as seen from this rank: `ptr` is the address of
    element [0, 0] in this process (the root's own pointer on the root, a CUDA IPC mapping opened from this rank's
    device elsewhere), `ld` the row pitch in elements."""

    def __init__(self, ptr: int, channels: int, samples: int, ld: int, mapping: Optional[int], keep=None):
        self.ptr, self.channels, self.samples, self.ld = ptr, channels, samples, ld
        self._mapping, self._keep = mapping, keep

    def rows(self, begin: int) -> int:
        return self.ptr + 4 * self.ld * begin

    def close(self) -> None:
        if self._mapping is not None:
            _cudart().cudaIpcCloseMemHandle(self._mapping)
            self._mapping = None


def _cudart():
    try:
        from cuda.bindings import runtime as cudart
    except ImportError:                                  # older cuda-python
        from cuda import cudart
    return cudart


def _cu_check(ret, what: str):
    err, rest = ret[0], ret[1:]
    if int(err) != 0:
        raise RuntimeError(f"{what}: {err}")
    return rest[0] if len(rest) == 1 else rest


def share_from_root(tensor, root: int = 0, group=None) -> PeerBlock:
    """Every rank gets a PeerBlock onto `tensor` of rank `root` (a planar fp32 CUDA tensor with unit inner stride;
    pass None elsewhere).  The root exports the allocation that holds the tensor as a CUDA IPC handle; every other
    rank opens it FROM ITS OWN DEVICE with cudaIpcMemLazyEnablePeerAccess (the way NCCL's P2P transport does), so
    the mapping is valid for kernels -- and TMA tensor maps -- of this rank's GPU, and loads / stores through it
    travel over NVLink / NVSwitch.  The root must keep `tensor` alive until every rank has closed its block."""
    import torch
    import torch.distributed as dist
    cudart = _cudart()
    rank = dist.get_rank(group)
    box = [None]
    if rank == root:
        if not (tensor.is_cuda and tensor.dtype == torch.float32 and tensor.dim() == 2 and tensor.stride(1) == 1):
            raise TypeError("share_from_root needs a planar fp32 CUDA tensor with unit inner stride")
        try:
            from cuda.bindings import driver as cu
        except ImportError:
            from cuda import cuda as cu
        base, _size = _cu_check(cu.cuMemGetAddressRange(tensor.data_ptr()), "cuMemGetAddressRange")
        handle = _cu_check(cudart.cudaIpcGetMemHandle(int(base)), "cudaIpcGetMemHandle")
        box[0] = (bytes(handle.reserved), tensor.data_ptr() - int(base), tensor.shape[0], tensor.shape[1],
                  tensor.stride(0), tensor.device.index)
    dist.broadcast_object_list(box, src=root, group=group)
    raw, offset, channels, samples, ld, owner = box[0]
    if rank == root:
        return PeerBlock(tensor.data_ptr(), channels, samples, ld, None, keep=tensor)
    cur = torch.cuda.current_device()
    if owner != cur and not torch.cuda.can_device_access_peer(cur, owner):
        raise RuntimeError(f"device {cur} cannot access device {owner} as a peer (no NVLink / PCIe P2P path)")
    _cu_check(cudart.cudaSetDevice(cur), "cudaSetDevice")
    handle = cudart.cudaIpcMemHandle_t()
    handle.reserved = raw
    mapped = _cu_check(cudart.cudaIpcOpenMemHandle(handle, cudart.cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")
    return PeerBlock(int(mapped) + offset, channels, samples, ld, int(mapped))


def process_on_root_block(plan, x_root: PeerBlock, y_root: PeerBlock, group=None) -> None:
    """Scatter + evaluate + gather in one kernel per rank: `x_root` / `y_root` are this rank's views (see
    share_from_root) of the planar blocks that live on the root; `plan` is this rank's plan for
    channel_range(channels, world, rank) of a one-input one-output graph.  The kernel's TMA loads pull the rank's
    rows out of the root's memory and its TMA stores put the results back there.  Returns after the local stream has
    been synchronised and every rank has arrived (the root may then read its output block)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = channel_range(x_root.channels, world, rank)
    if e - b != plan.channels:
        raise ValueError(f"plan has {plan.channels} channels, this rank owns {e - b}")
    if e > b:
        stream = torch.cuda.current_stream().cuda_stream
        plan.process_ptrs([x_root.rows(b)], [y_root.rows(b)], x_root.samples, x_root.ld, y_root.ld, stream)
    torch.cuda.synchronize()
    dist.barrier(group)
