#!/usr/bin/env python
"""HBM bandwidth probes with library kernels (torch), for the write-only / read-write rooflines quoted in
DESIGN.md: copy (read + write), fill (write only), sum (read only), 4 GiB each, best of 10, CUDA events."""
import json, torch
N = 1 << 30
a = torch.empty(N, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
def best(fn, nbytes):
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return round(nbytes / min(ts) / 1e6, 1)
a.fill_(1.0); b.fill_(0.0); torch.cuda.synchronize()
print(json.dumps({"copy_gbs": best(lambda: b.copy_(a), 8 * N), "fill_gbs": best(lambda: b.fill_(2.0), 4 * N),
                  "zero_gbs": best(lambda: b.zero_(), 4 * N), "sum_gbs": best(lambda: a.sum(), 4 * N)}))
