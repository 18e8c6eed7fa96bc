#!/usr/bin/env python
"""Multi-GPU edge step on real GPUs (run under torchrun, NCCL): a block that lives on rank 0 is scattered by
channel ranges, every rank evaluates its shard with its own plan, the outputs are gathered, and rank 0
compares with the same block evaluated on one GPU -- bit for bit (channels are independent).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/shard_check.py
"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import zignal_b200 as zg
from zignal_b200 import workloads as fo
from zignal_b200.shard import channel_range, scatter_channels, gather_channels


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C, T = 8192 + 37, 4096                     # uneven split on purpose
    g = zg.compile(fo.biquad_cascade(4))
    full = None
    if rank == 0:
        gen = torch.Generator(device=dev).manual_seed(5)
        full = torch.rand((C, T), generator=gen, device=dev) * 2 - 1
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    own = scatter_channels(full, C, T, root=0, device=dev)
    b, e = channel_range(C, world, rank)
    y = g.plan(channels=e - b, device=local).process([own])[0]
    out = gather_channels(y.contiguous(), C, T, root=0)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    ref = None
    if rank == 0:
        ref = g.plan(channels=C, device=local).process([full])[0]
        torch.cuda.synchronize()
        print(json.dumps({"check": "scatter -> shard plans -> gather == one GPU", "world": world, "channels": C, "samples": T,
                          "bit_identical": bool(torch.equal(out, ref)), "seconds_incl_plan_creation": round(dt, 3)}), flush=True)
    dist.barrier()

    # ---- the same edge step fused into the kernels: every rank streams its rows from / to the root's HBM over
    #      NVLink peer memory (TMA tensor maps on IPC-mapped buffers), larger block, both paths timed ----
    from zignal_b200.shard import share_from_root, process_on_root_block
    C2, T2 = 65536, 8192
    x_root = y_root = None
    if rank == 0:
        gen = torch.Generator(device=dev).manual_seed(6)
        x_root = torch.rand((C2, T2), generator=gen, device=dev) * 2 - 1
        y_root = torch.empty_like(x_root)
    xa = share_from_root(x_root, 0)
    ya = share_from_root(y_root, 0)
    if rank != 0:                                   # the mapping itself, before any kernel dereferences it
        probe = torch.empty(16, device=dev)
        from zignal_b200.shard import _cudart
        err = _cudart().cudaMemcpy(probe.data_ptr(), xa.rows(b2 := channel_range(C2, world, rank)[0]), 64, _cudart().cudaMemcpyKind.cudaMemcpyDefault)[0]
        print(json.dumps({"rank": rank, "peer_mapping_memcpy": str(err), "first": probe[:2].tolist()}), flush=True)
    b2, e2 = channel_range(C2, world, rank)
    plan = g.plan(channels=e2 - b2, device=local)
    plan_n = g.plan(channels=e2 - b2, device=local)
    times = {}
    for name in ("nccl", "peer", "nccl", "peer"):
        plan.reset(); plan_n.reset()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "peer":
            process_on_root_block(plan, xa, ya)
        else:
            own = scatter_channels(x_root, C2, T2, root=0, device=dev)
            yo = plan_n.process([own])[0]
            out_n = gather_channels(yo.contiguous(), C2, T2, root=0)
            torch.cuda.synchronize(); dist.barrier()
        times[name] = time.perf_counter() - t0
    if rank == 0:
        ref2 = g.plan(channels=C2, device=local).process([x_root])[0]
        torch.cuda.synchronize()
        print(json.dumps({"check": "peer-memory edge step (kernels read/write the root's HBM over NVLink) == one GPU",
                          "world": world, "channels": C2, "samples": T2, "bit_identical": bool(torch.equal(y_root, ref2)),
                          "nccl_also_identical": bool(torch.equal(out_n, ref2)),
                          "ms_peer_fused": round(times["peer"] * 1e3, 2), "ms_nccl_scatter_compute_gather": round(times["nccl"] * 1e3, 2),
                          "GBps_through_root_peer": round(2 * 4 * C2 * T2 * (world - 1) / world / times["peer"] / 1e9, 1)}), flush=True)
    dist.barrier()
    xa.close(); ya.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
