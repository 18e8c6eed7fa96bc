#!/usr/bin/env python
"""Tuning sweep for the FIR kernel K3 (GPU box only): kernel-only Msamples/s for combinations of
mode, warps per CTA (ZG_TUNE_WPC) and time segments (ZG_TUNE_SEGS).  One JSON line per point.

    python tools/sweep_fir.py [--channels 32768] [--samples 8192] [--taps 256] [--points "mode=exact,fast;wpc=8,12;segs=0"]
"""
import argparse, itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import zignal_b200 as zg
from zignal_b200 import workloads as fo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=32768)
    ap.add_argument("--samples", type=int, default=8192)
    ap.add_argument("--taps", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--points", default="mode=exact,fast;wpc=0;segs=0")
    a = ap.parse_args()
    axes = {}
    for part in a.points.split(";"):
        k, v = part.split("=")
        axes[k] = v.split(",")
    for k, d in (("mode", ["exact"]), ("wpc", ["0"]), ("segs", ["0"])):
        axes.setdefault(k, d)
    C, T = a.channels, a.samples
    x = torch.rand((C, T), device="cuda") * 2 - 1
    y = torch.empty_like(x)
    g = zg.compile(fo.fir_expr(fo.fir_taps(a.taps)))
    keys = list(axes)
    for combo in itertools.product(*[axes[k] for k in keys]):
        pt = dict(zip(keys, combo))
        for env, k in (("ZG_TUNE_WPC", "wpc"), ("ZG_TUNE_SEGS", "segs")):
            if pt[k] != "0": os.environ[env] = pt[k]
            else: os.environ.pop(env, None)
        try:
            plan = g.plan(channels=C, mode=zg.MODE_EXACT if pt["mode"] == "exact" else zg.MODE_FAST)
            for _ in range(2): plan.process([x], [y])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters): plan.process([x], [y])
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            info = plan.info()
            pt.update(C=C, T=T, taps=a.taps, ms=round(ms, 4), msamples=round(C * T / ms / 1e3), gbs=round(8 * C * T / ms / 1e6),
                      tflops=round(2 * a.taps * C * T / ms / 1e9, 2), kernel=info.kernel.decode(),
                      threads=info.threads_per_cta, ring=info.stages, seg_boxes=info.boxes, smem=info.smem_bytes,
                      regs=info.regs_per_thread)
        except Exception as e:
            pt["error"] = str(e)[:300]
        print(json.dumps(pt), flush=True)


if __name__ == "__main__":
    main()
