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
    if rank == 0:
        ref = g.plan(channels=C, device=local).process([full])[0]
        torch.cuda.synchronize()
        print(json.dumps({"check": "scatter -> shard plans -> gather == one GPU", "world": world, "channels": C, "samples": T,
                          "bit_identical": bool(torch.equal(out, ref)), "seconds_incl_plan_creation": round(dt, 3)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
