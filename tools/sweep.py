#!/usr/bin/env python
"""Tuning sweep (GPU box only): kernel-only Msamples/s of the biquad cascade for combinations of the
ZG_TUNE_* launch-geometry overrides, modes, layouts and workloads.  Prints one JSON line per point.

    python tools/sweep.py [--workload ns] [--iters 20] [--points "mode=exact,fast;boxes=1,2;wpc=0;stages=0"]
"""
import argparse, itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import zignal_b200 as zg
from zignal_b200 import workloads as fo

WORK = {"c3": (65536, 16384), "c5": (131072, 4096), "ns": (65536, 8192), "c2": (4096, 65536), "mid": (16384, 16384), "c32k": (32768, 8192), "c8k": (8192, 32768)}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ns")
    ap.add_argument("--shape", default="", help="C,T instead of a named workload")
    ap.add_argument("--brief", action="store_true", help="print the axes, the kernel, ms and GB/s only")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sections", type=int, default=4)
    ap.add_argument("--pad", type=int, default=0, help="extra floats of row pitch (planar buffers)")
    ap.add_argument("--graph", default="biquad", choices=["biquad", "osc", "poly", "copy", "comb"],
                    help="osc = configs[2] (dirac in, fp32 out), poly = configs[4] (dirac in, bf16 out), copy = y = 1.0f*x")
    ap.add_argument("--points", default="mode=exact,fast;boxes=1,2;wpc=0;stages=0;layout=planar;coef=uniform")
    a = ap.parse_args()
    axes = {}
    for part in a.points.split(";"):
        k, v = part.split("=")
        axes[k] = v.split(",")
    for k, d in (("mode", ["exact"]), ("boxes", ["0"]), ("wpc", ["0"]), ("stages", ["0"]), ("layout", ["planar"]), ("coef", ["uniform"]), ("lanes", ["0"]), ("late", ["0"]), ("hint", ["0"]), ("promo", ["0"]), ("cu", ["0"]), ("jit", ["0"]), ("pf", ["0"]), ("pfd", ["0"]), ("tp", ["0"]), ("segs", ["0"]), ("warm", ["0"]), ("split", ["0"]), ("sg", ["0"]), ("spw", ["0"]), ("hb", ["0"])):
        axes.setdefault(k, d)
    C, T = [int(v) for v in a.shape.split(",")] if a.shape else WORK[a.workload]
    x = torch.empty((C, T + a.pad), device="cuda")[:, :T]
    x.copy_(torch.rand((C, T), device="cuda") * 2 - 1)
    y = torch.empty((C, T + a.pad), device="cuda")[:, :T]
    xi, yi = None, None
    keys = list(axes)
    for combo in itertools.product(*[axes[k] for k in keys]):
        pt = dict(zip(keys, combo))
        for env, k in (("ZG_TUNE_BOXES", "boxes"), ("ZG_TUNE_WPC", "wpc"), ("ZG_TUNE_STAGES", "stages"), ("ZG_TUNE_LATE_REFILL", "late"), ("ZG_TUNE_L2HINT", "hint"), ("ZG_TUNE_L2PROMO", "promo"), ("ZG_TUNE_CHUNK_UNROLL", "cu"), ("ZG_TUNE_PF", "pf"), ("ZG_TUNE_PFD", "pfd"), ("ZG_TUNE_TP", "tp"), ("ZG_TUNE_SEGS", "segs"), ("ZG_TUNE_WARM", "warm"), ("ZG_TUNE_SPLIT", "split"), ("ZG_TUNE_SPLIT_G", "sg"), ("ZG_TUNE_SPLIT_SPW", "spw"), ("ZG_TUNE_SPLIT_HB", "hb")):
            if pt[k] != "0": os.environ[env] = pt[k]
            else: os.environ.pop(env, None)
        inter = pt["layout"] == "interleaved"
        if inter and xi is None:
            xi = x.t().contiguous(); yi = torch.empty_like(xi)
        extra = {}
        out_dt, bytes_per_sample, has_in = torch.float32, 8, True
        if a.graph == "osc":
            g = zg.compile(fo.osc_lp_expr()); extra = dict(input_kind=[zg.IN_DIRAC]); bytes_per_sample, has_in = 4, False
        elif a.graph == "poly":
            g = zg.compile(fo.poly_voice_expr()); extra = dict(input_kind=[zg.IN_DIRAC], io_dtype=zg.BF16)
            out_dt, bytes_per_sample, has_in = torch.bfloat16, 2, False
        elif a.graph == "copy":
            g = zg.compile("0x1p+0f*_1")
        elif a.graph == "comb":                        # feedback comb + feed-forward echo: two long delay lines (rings in HBM)
            g = zg.compile("~(_2 + 0.5f*_1[_441]) |= (_1 + 0.25f*_1[_1000])")
        elif pt["coef"] == "uniform":
            g = zg.compile(fo.biquad_cascade(a.sections))
        else:
            g = zg.compile(fo.biquad_cascade_params(a.sections))
        try:
            plan = g.plan(channels=C, mode=zg.MODE_EXACT if pt["mode"] == "exact" else zg.MODE_FAST,
                          layout=zg.INTERLEAVED if inter else zg.PLANAR, lanes_per_channel=int(pt["lanes"]), force_jit=pt["jit"] == "1", **extra)
            if pt["coef"] != "uniform":
                import numpy as np
                for k in range(a.sections):
                    co = fo.rbj_lowpass(440.0 * 2 ** k)
                    for j in range(5):
                        plan.set_param(5 * k + j, np.full(C, co[j], np.float32))
            bi, bo = ([xi], [yi]) if inter else ([x], [y])
            if out_dt != torch.float32: bo = [torch.empty(bo[0].shape, dtype=out_dt, device="cuda")]
            if not has_in: bi = [None]
            _proc = plan.process
            plan.process = lambda i, o: _proc(i, o, n_samples=T)
            for _ in range(3): plan.process(bi, bo)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters): plan.process(bi, bo)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            info = plan.info()
            pt.update(kernel=info.kernel.decode(), pad=a.pad, ms=round(ms, 4), msamples=round(C * T / ms / 1e3), gbs=round(bytes_per_sample * C * T / ms / 1e6),
                      threads=info.threads_per_cta, stages_used=info.stages, boxes_used=info.boxes, smem=info.smem_bytes, lanes_used=info.lanes_per_channel, segs_used=info.time_segments, seg_len=info.segment_samples, warm_used=info.warmup_samples,
                      regs=info.regs_per_thread)
        except Exception as e:
            pt["error"] = str(e)[:200]
        pt = {k: v for k, v in pt.items() if v != "0" or k not in axes}        # axes left at their default are not printed
        if a.brief:
            pt = {k: v for k, v in pt.items() if k in axes or k in ("kernel", "ms", "gbs", "threads", "stages_used", "boxes_used", "error")}
        print(json.dumps(pt), flush=True)

if __name__ == "__main__":
    main()
