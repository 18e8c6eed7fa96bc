#!/usr/bin/env python
"""bench.py -- Msamples/s of the N-channel biquad chain on 1..8 B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ns|c2] [--coef uniform|per-channel]
    python bench.py --impl reference ...      # the reference's own CPU loop on the host cores
    (N > 1: launched by torch.distributed.run, one rank per GPU)

A step = one pass of the hot path (one zg_process() = ONE kernel launch per rank) over one block of
synthetic input: `ns` = 65 536 channels x 8192 samples per GPU (the north star's target shape,
SURVEY.md 8d), `c2` = 4096 x 65 536 (BASELINE configs[1]); 4 direct-form-1 biquad sections in series,
the reference's own graph spelling (test/benchmark.cpp:25-33), stable RBJ low-pass coefficients.
Channels are independent, so N GPUs = N x the channels (weak scaling), no data-path collective.

  value     whole-job Msamples/s, inputs resident in HBM, CUDA events on the launching stream
  e2e       same metric through zg_process_host(): pinned host buffers, H2D + kernel + D2H timed
  roofline  algorithmic bytes (8 B per sample: 4 in + 4 out) / average launch time vs measured HBM peak
  cpu_baseline  the reference's hand-written biquad loop (oracle/_ref, all host threads), bounded sample; this
            CPU leg is also the only place the arm touches oracle/ (a spot check of the timed kernel's output
            against the oracle rides along as cpu_baseline.parity_spot_check); graphs and coefficients of
            the GPU legs come from zignal_b200/workloads.py
  also      the other biquad shape (c2 when the headline is ns and vice versa), device-resident, same rules;
            c2_scan = configs[1] in FAST mode, cut in time (zg_plan_opts.time_parallel);
            ns_per_channel = the north-star shape with the per-channel coefficients SURVEY.md 8(d) prescribes;
            ns_k1 = the north-star shape on the lane-per-channel kernel (zg_plan_opts.section_warps = 1)
  edge      (N > 1) a 65 536 x 8192 block resident on rank 0 -> N shard plans -> back on rank 0, three ways:
            NCCL scatter + compute + gather, kernels working on the root's block over NVLink peer memory, compute only
Default mode is EXACT: bit-identical to the reference's x86 build, checked on sampled channels.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SECTIONS = 4
WORKLOADS = {
    # name: (channels per GPU, samples per block)
    "ns": (65536, 8192),
    "c2": (4096, 65536),
}
BYTES_PER_SAMPLE = 8          # fp32 in + fp32 out; state/coefficients amortise to < 0.1 %


def workload_string(name, layout="planar"):
    C, T = WORKLOADS[name]
    return (f"{name}: {C} channels x {T} samples per GPU x {SECTIONS} DF1 biquad sections "
            f"(flowz `fwd |= bwd` x{SECTIONS}), fp32 {layout}")


def _peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of "
                  "the same kernel and shape (profiles/traffic.json), not from this run")


def _traffic(workload):
    """dram bytes per launch from the committed `ncu --set full` capture of the same kernel, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons during the timed region (NVML, ~2 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- the reference arm / cpu baseline: the reference's own loop on the host cores ---------------------

def _ref_lib():
    """oracle/_ref/libzg_ref_custom.so (the reference's benchmark.cpp compiled where it lies), on ALL host threads:
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would silently time one core."""
    so = os.path.join(ROOT, "oracle", "_ref", "libzg_ref_custom.so")
    if os.path.exists(so):
        lib = ctypes.CDLL(so)
        try:
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count() or 1)
        except OSError:
            pass
        return lib
    return None


def cpu_reference_rate(target_seconds=12.0, samples=8192):
    """Times `SECTIONS` hand-written DF1 biquads in series per channel (the reference's make_custom,
    test/benchmark.cpp:35-47, compiled where it lies into oracle/_ref) on all host threads, on a bounded
    sample of the workload.  Falls back to the oracle's emitted C tick ("port") if _ref is absent."""
    import numpy as np
    import flowz_oracle as fo
    cores = os.cpu_count() or 1
    lib = _ref_lib()
    P = ctypes.POINTER(ctypes.c_float)

    def run_ref(C):
        x = fo.noise(C, samples, seed=0) * np.float32(0.01)
        y = np.empty_like(x)
        y.fill(1.0)                               # pages touched before the clock starts (np.zeros would map them lazily)
        t0 = time.perf_counter()
        rc = lib.zg_ref_df1_chain(SECTIONS, x.ctypes.data_as(P), y.ctypes.data_as(P), ctypes.c_long(C), ctypes.c_long(samples))
        dt = time.perf_counter() - t0
        assert rc == 0
        return dt

    port = None

    def run_port(C):
        nonlocal port
        x = fo.noise(C, samples, seed=0)
        port = fo.COracle(fo.biquad_cascade(SECTIONS), C)
        t0 = time.perf_counter()
        port.process([x])
        return time.perf_counter() - t0

    run, kind = (run_ref, "reference") if lib is not None else (run_port, "port")
    C = 64 * cores
    dt = run(C)                                   # calibration (also warms the threads up)
    rate = C * samples / dt
    C = int(max(cores, min(rate * target_seconds / samples, 262144)))
    dt = run(C)
    return {"value": C * samples / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "sample": f"{C} channels x {samples} samples x {SECTIONS} DF1 sections, {dt:.1f} s, "
                      f"{'reference make_custom loop (coefficients of test/benchmark.cpp:18-23)' if kind == 'reference' else 'oracle C tick'}, "
                      f"OpenMP over channels"}, (C, dt)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C_w, T_w = WORKLOADS[args.workload]
    # each step = a bounded sample of the workload; size it once so the whole run ends in minutes
    base, (C, dt) = cpu_reference_rate(target_seconds=2.0, samples=min(T_w, 8192))
    import numpy as np
    import flowz_oracle as fo
    lib = _ref_lib()
    P = ctypes.POINTER(ctypes.c_float)
    samples = min(T_w, 8192)
    x = fo.noise(C, samples, seed=0) * np.float32(0.01)
    y = np.empty_like(x)

    def step():
        if lib is not None:
            lib.zg_ref_df1_chain(SECTIONS, x.ctypes.data_as(P), y.ctypes.data_as(P), ctypes.c_long(C), ctypes.c_long(samples))
        else:
            fo.COracle(fo.biquad_cascade(SECTIONS), C).process([x])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = C * samples * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s, N-channel 4-section biquad chain", "value": value,
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload),
                   "step_sample": f"{C} channels x {samples} samples per step on the host"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": base["cores"], "kind": base["kind"],
                         "sample": base["sample"]},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- the other BASELINE configs (parity-tested in tests/; timed here for the record, not the headline) ----

def other_configs(zg, wl, torch, dev, local, world, rank, args, barrier, dist):
    """configs[2] osc >> one-pole LP (65 536 voices, source-only: 4 B/sample written), configs[3] 256-tap FIR
    x 32 768 channels (8 B/sample, but bound by FP32 issue: 511 instructions per sample in EXACT mode),
    configs[4] polyphonic chain, bf16 storage, 131 072 voices per GPU (2 B/sample written).  Same timing
    rules as the headline: warm-up, CUDA events on the launching stream, max over ranks, outputs > L2."""
    mode = zg.MODE_EXACT if args.mode == "exact" else zg.MODE_FAST
    peak = _peak_hbm()[0]
    steps = max(3, min(args.steps, 10))
    out = {}
    cases = [
        ("c3", "configs[2]: 65536-voice sine osc >> one-pole LP, dirac-excited, fp32 out", wl.osc_lp_expr(), 65536, 16384,
         dict(input_kind=[zg.IN_DIRAC]), None, torch.float32, 4, None),
        ("c4", "configs[3]: 256-tap FIR x 32768 channels x 8192 samples, fp32", wl.fir_expr(wl.fir_taps(256)), 32768, 8192,
         dict(), torch.float32, torch.float32, 8, 511 if args.mode == "exact" else 256),
        ("c5", "configs[4]: polyphonic chain osc >> biquad >> (biquad ~ feedback), 131072 voices per GPU x 4096 samples, "
               "bf16 out, dirac-excited", wl.poly_voice_expr(), 131072, 4096,
         dict(input_kind=[zg.IN_DIRAC], io_dtype=zg.BF16), None, torch.bfloat16, 2, None),
    ]
    cases.append(("c4_tc", "configs[3] in FAST mode: the Toeplitz contraction on tcgen05 tensor cores (3xTF32), 256-tap FIR x 32768 "
                           "channels x 8192 samples, fp32", wl.fir_expr(wl.fir_taps(256)), 32768, 8192, dict(mode=zg.MODE_FAST),
                  torch.float32, torch.float32, 8, None))
    for name, desc, expr, C, T, kw, in_dt, out_dt, bytes_per_sample, instr_per_sample in cases:
        try:
            kw = dict(kw)
            plan = zg.compile(expr).plan(channels=C, device=local, mode=kw.pop("mode", mode), **kw)
            x = (torch.rand((C, T), device=dev, dtype=torch.float32) * 2 - 1).to(in_dt) if in_dt is not None else None
            y = torch.empty((C, T), device=dev, dtype=out_dt)
            for _ in range(3):
                plan.process([x], [y], n_samples=T)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                plan.process([x], [y], n_samples=T)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1) / steps
            if dist is not None:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            info = plan.info()
            ent = {"workload": desc, "value": world * C * T / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms,
                   "steps": steps, "bytes_per_sample": bytes_per_sample,
                   "roofline_frac": bytes_per_sample * C * T / (ms * 1e-3) / 1e9 / peak, "kernel": info.kernel.decode(),
                   "jit": info.jit, "regs": info.regs_per_thread, "mode": args.mode}
            if name == "c5":
                ent["bound"] = "FP32 issue (21 arithmetic instructions per 2-byte sample after delayed-product reuse), not HBM"
            if name == "c4_tc":
                # tensor roofline: algorithmic flops = 2 x 256 taps per sample; executed = 3 operand products x 288/256
                # band padding.  The TF32 peak is not in MEASURED_PEAKS.json: measured here with a cuBLAS TF32 GEMM.
                ent["mode"] = "fast"
                tf32 = _tf32_peak(torch, dev)
                algo = 2 * 256 * C * T / (ms * 1e-3) / 1e12
                ent["roofline"] = {"bound": "tensor", "achieved": algo, "unit": "TFLOP/s", "peak": tf32,
                                   "frac": algo / tf32 if tf32 else None,
                                   "peak_source": "measured here: torch.matmul fp32 8192^3 with TF32 enabled (cuBLAS), best of 5",
                                   "executed_tflops": algo * 3 * 288 / 256, "traffic": _traffic("c4_tc"),
                                   "executed_frac": algo * 3 * 288 / 256 / tf32 if tf32 else None,
                                   "note": "3xTF32 (hi*hi + lo*hi + hi*lo) and a 288-row band per 256 taps: 3.375 tensor flops "
                                           "per algorithmic flop; executed_frac is against the cuBLAS TF32 GEMM timed here "
                                           "(profiles/r02_c4_fir_tc_ncu.txt)"}
                ent["numerics"] = "<= 1e-5 block-relative against the oracle (tests/test_fir_tc.py; measured 3e-6: the tensor core truncates its fp32 accumulation)"
            if instr_per_sample:
                sm = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = ClockSampler(local).max_mhz or 1965
                ent["fp32_issue_frac"] = instr_per_sample * C * T / (ms * 1e-3) / (sm * 128 * mhz * 1e6)
                ent["bound"] = f"FP32 issue ({instr_per_sample} instructions per sample), not HBM"
            out[name] = ent
            del plan, x, y
            torch.cuda.empty_cache()
        except Exception as e:                      # the headline must still be reported
            out[name] = {"workload": desc, "error": str(e)[:300]}
    return out


# ---- the edge step: the only exchange the path has, when a block lives on one GPU ---------------

def edge_step(zg, wl, torch, dist, dev, local, world, rank, args):
    """A planar 65 536 x 8192 fp32 block resident on rank 0 is evaluated by `world` shard plans and ends up on rank 0
    again, three ways, each timed on the device (CUDA events per rank, max over ranks) after a warm-up round:
      nccl          one grouped NCCL send/recv scatter, the shard kernels, one grouped gather (shard.scatter_channels /
                    gather_channels) -- the literal reading of `a single NCCL scatter/gather at the edges`;
      peer_fused    no collective: every rank's TMA tensor maps point at the root's HBM (CUDA IPC peer mappings,
                    shard.share_from_root), its kernel pulls its rows over NVLink / NVSwitch and stores the results back,
                    tile by tile, overlapped with the arithmetic (shard.process_on_root_block);
      compute_only  every rank on a local copy of its shard: what the exchange costs on top.
    The root's NVLink port (900 GB/s per direction) carries (world - 1) / world of the block out and the same back."""
    from zignal_b200.shard import (channel_range, gather_channels, process_on_root_block, scatter_channels,
                                   share_from_root)
    C, T = WORKLOADS["ns"]
    mode = zg.MODE_EXACT if args.mode == "exact" else zg.MODE_FAST
    g = zg.compile(wl.biquad_cascade(SECTIONS))
    b, e = channel_range(C, world, rank)
    x_root = y_root = None
    if rank == 0:
        gen = torch.Generator(device=dev).manual_seed(99)
        x_root = torch.rand((C, T), generator=gen, device=dev) * 2 - 1
        y_root = torch.empty_like(x_root)
    xa, ya = share_from_root(x_root, 0), share_from_root(y_root, 0)
    plan = g.plan(channels=e - b, device=local, mode=mode)
    own = scatter_channels(x_root, C, T, root=0, device=dev)           # also the compute-only input
    y_own = torch.empty_like(own)
    out_nccl = torch.empty_like(x_root) if rank == 0 else None
    steps = 5

    def timed(fn):
        fn()                                                           # warm-up (NCCL channels, peer mappings, L2 state)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_nccl():
        o = scatter_channels(x_root, C, T, root=0, device=dev)
        plan.process([o], [y_own])
        gather_channels(y_own, C, T, root=0, out=out_nccl)

    def run_peer():
        stream = torch.cuda.current_stream().cuda_stream
        plan.process_ptrs([xa.rows(b)], [ya.rows(b)], T, xa.ld, ya.ld, stream)

    def run_local():
        plan.process([own], [y_own])

    ms_nccl = timed(run_nccl)
    ms_peer = timed(run_peer)
    ms_local = timed(run_local)
    # every rank's kernel must have finished before the root reads its block
    plan.reset(); process_on_root_block(plan, xa, ya)
    plan.reset(); run_nccl(); torch.cuda.synchronize(); dist.barrier()
    res = None
    if rank == 0:
        ref = g.plan(channels=C, device=local, mode=mode).process([x_root])[0]
        torch.cuda.synchronize()
        port_bytes = 4.0 * C * T * (world - 1) / world                # per direction
        res = {"block": f"{C} channels x {T} samples fp32 planar, resident on rank 0 (2 GiB in, 2 GiB out)",
               "steps": steps, "timing": "CUDA events per rank, max over ranks",
               "nccl_scatter_compute_gather_ms": ms_nccl, "peer_fused_ms": ms_peer, "compute_only_ms": ms_local,
               "bit_identical": {"peer_fused": bool(torch.equal(y_root, ref)), "nccl": bool(torch.equal(out_nccl, ref))},
               "root_port_gbs_per_direction": {"peer_fused": port_bytes / (ms_peer * 1e-3) / 1e9,
                                               "nccl": port_bytes / (ms_nccl * 1e-3) / 1e9},
               "root_port_frac_of_900": {"peer_fused": port_bytes / (ms_peer * 1e-3) / 1e9 / 900.0,
                                         "nccl": port_bytes / (ms_nccl * 1e-3) / 1e9 / 900.0},
               "msamples_per_s": {"peer_fused": C * T / (ms_peer * 1e-3) / 1e6, "nccl": C * T / (ms_nccl * 1e-3) / 1e6,
                                  "compute_only": C * T / (ms_local * 1e-3) / 1e6},
               "limiter": "the root's NVLink port: (N-1)/N of the block leaves and returns through it; both directions "
                          "run at once in the fused form (loads and stores of the same kernels), one after the other "
                          "(scatter, then gather) through NCCL"}
        del ref
    dist.barrier()
    xa.close(); ya.close()
    dist.barrier()
    return res


def _tf32_peak(torch, dev, n=8192):
    """Dense TF32 tensor throughput of this GPU through cuBLAS (TFLOP/s), for the roofline of the tensor-core FIR."""
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((n, n), device=dev)
        b = torch.randn((n, n), device=dev)
        best = 0.0
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c = a @ b
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b, c
        return best
    except Exception:
        return None


def _bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, before any pinned host buffer is allocated:
    with one rank per GPU on a two-socket box, host staging memory that sits on the other socket turns the e2e
    path (H2D + D2H of every block) into cross-socket traffic.  Best effort; silently skipped if NVML or the
    affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# ---- our arm ---------------------------------------------------------------------------------------------

def run_ours(args):
    import numpy as np
    import torch
    import zignal_b200 as zg
    from zignal_b200 import workloads as wl      # graph text + coefficients; nothing under oracle/ is touched by this arm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_rate(workload, coef=None, mode_name=None, **plan_kw):
        """K launches of one zg_process() each between CUDA events; returns the plan, buffers and timing."""
        C, T = WORKLOADS[workload]
        coef = coef or args.coef
        mode_name = mode_name or args.mode
        if coef == "uniform":
            graph = zg.compile(wl.biquad_cascade(SECTIONS))
            params = []
        else:
            graph = zg.compile(wl.biquad_cascade_params(SECTIONS))
            c = np.arange(C, dtype=np.float64) + rank * C
            params = []
            for k in range(SECTIONS):
                per = np.array([wl.rbj_lowpass(440.0 * 2 ** k * (1.0 + ci / (C * world))) for ci in c], np.float32)
                params += [per[:, j].copy() for j in range(5)]
        mode = zg.MODE_EXACT if mode_name == "exact" else zg.MODE_FAST
        plan = graph.plan(channels=C, device=local, mode=mode,
                          layout=zg.INTERLEAVED if args.layout == "interleaved" else zg.PLANAR, **plan_kw)
        for i, p in enumerate(params):
            plan.set_param(i, p)
        shape = (T, C) if args.layout == "interleaved" else (C, T)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        x = torch.empty(shape, device=dev, dtype=torch.float32).uniform_(-1.0, 1.0, generator=gen)     # U(-1, 1), in place
        y = torch.empty_like(x)
        for _ in range(max(args.warmup, 3)):
            plan.process([x], [y])
        barrier()
        l0 = plan.info().launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            barrier()
            e0.record()
            for _ in range(args.steps):
                plan.process([x], [y])
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        launches = plan.info().launches - l0
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return plan, x, y, shape, ms / args.steps, launches, clk

    C, T = WORKLOADS[args.workload]
    plan, x, y, shape, ms_per_step, launches, clk = device_rate(args.workload)
    value = world * C * T / (ms_per_step * 1e-3) / 1e6
    info = plan.info()

    # ---- e2e: host buffers through zg_process_host (H2D + kernel + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(shape, dtype=torch.float32).pin_memory()
        hy = torch.empty(shape, dtype=torch.float32).pin_memory()
        hx.copy_(x)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        plan.process_host([hx], [hy])                       # warm-up (allocates the device staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.process_host([hx], [hy])
            _ = float(hy[0, 0])                              # the result is read on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * C * T * e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(C * T * 4), "d2h_bytes_per_step": int(C * T * 4), "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3, "api": "zg_process_host (pinned host buffers)",
               "host_chunks": plan.info().host_chunks,
               "gbs_per_rank_both_directions": 2 * C * T * 4 * e2e_steps / dt / 1e9}
        # what the host link gives this rank with all ranks copying at once and no kernel in the way: one direction
        # alone, then both together -- tells PCIe (per-GPU link) from host DRAM / root-complex saturation
        copies = {}
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        for name, do_h2d, do_d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
            barrier()
            t0 = time.perf_counter()
            if do_h2d:
                with torch.cuda.stream(s1):
                    x.copy_(hx, non_blocking=True)
            if do_d2h:
                with torch.cuda.stream(s2):
                    hy.copy_(y, non_blocking=True)
            torch.cuda.synchronize()
            dtc = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dtc], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtc = float(t.item())
            copies[name + "_gbs_per_rank"] = (int(do_h2d) + int(do_d2h)) * C * T * 4 / dtc / 1e9
        e2e["copy_only"] = copies
        e2e["limiter"] = ("host link: a step moves 2 x %.2f GB per rank; the kernel is %.1f %% of the step" %
                          (C * T * 4 / 1e9, 100 * ms_per_step / (dt / e2e_steps * 1e3)))
        del hx, hy
        # the one lever on a link-bound step: bf16 sample storage halves the bytes (state, coefficients and arithmetic stay
        # fp32; outputs rounded to nearest even -- bit-identical to the oracle fed the bf16-rounded block,
        # tests/test_gpu_parity.py).  Reported beside the fp32 number, never instead of it.
        try:
            plan_h = zg.compile(wl.biquad_cascade(SECTIONS)).plan(
                channels=C, device=local, mode=zg.MODE_EXACT if args.mode == "exact" else zg.MODE_FAST,
                layout=zg.INTERLEAVED if args.layout == "interleaved" else zg.PLANAR, io_dtype=zg.BF16)
            hxb = torch.empty(shape, dtype=torch.bfloat16).pin_memory()
            hyb = torch.empty(shape, dtype=torch.bfloat16).pin_memory()
            hxb.copy_(x)
            plan_h.process_host([hxb], [hyb])
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                plan_h.process_host([hxb], [hyb])
                _ = float(hyb[0, 0])
            torch.cuda.synchronize()
            dtb = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dtb], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtb = float(t.item())
            e2e["bf16_io"] = {"value": world * C * T * e2e_steps / dtb / 1e6, "unit": "Msamples/s",
                              "h2d_bytes_per_step": int(C * T * 2), "d2h_bytes_per_step": int(C * T * 2),
                              "ms_per_step": dtb / e2e_steps * 1e3, "kernel": plan_h.info().kernel.decode(),
                              "gbs_per_rank_both_directions": 2 * C * T * 2 * e2e_steps / dtb / 1e9,
                              "note": "zg_plan_opts.io_dtype = ZG_BF16: samples cross the link and live in HBM as bf16"}
            del plan_h, hxb, hyb
        except Exception as e:
            e2e["bf16_io"] = {"error": str(e)[:300]}

    # ---- CPU leg, part 1 (rank 0, N = 1 only; the one place this arm may execute oracle/): a spot check of the
    #      timed kernel's output against the oracle -- the checker, not the thing measured ----
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu and args.coef == "uniform" and args.layout == "planar":
        import flowz_oracle as fo
        plan.reset()                                         # the oracle starts from zero state too
        plan.process([x], [y])
        torch.cuda.synchronize()
        idx = [0, 1, C // 2, C - 1]
        ref = fo.COracle(fo.biquad_cascade(SECTIONS), len(idx)).process([x[idx].cpu().numpy()])[0]
        got = y[idx].cpu().numpy()
        den = np.abs(ref).max(axis=1)
        parity = {"max_block_rel_err": float((np.abs(got.astype(np.float64) - ref).max(axis=1) / den).max()),
                  "bit_identical": bool(np.array_equal(got, ref)), "channels_checked": idx}

    # ---- the other biquad configuration (BASELINE configs[1] vs the north-star shape), same timing rules ----
    also = None
    if not args.no_also:
        other = "c2" if args.workload == "ns" else "ns"
        del x, y
        torch.cuda.empty_cache()
        Co, To = WORKLOADS[other]
        plan_o, xo, yo, _, ms_o, _, _ = device_rate(other)
        io = plan_o.info()
        also = {other: {"workload": workload_string(other, args.layout),
                        "value": world * Co * To / (ms_o * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_o,
                        "roofline_frac": BYTES_PER_SAMPLE * Co * To / (ms_o * 1e-3) / 1e9 / _peak_hbm()[0],
                        "kernel": io.kernel.decode(), "lanes_per_channel": io.lanes_per_channel,
                        "traffic": _traffic(other)}}
        if other == "c2" and args.mode == "exact":
            also[other]["bound"] = ("recurrence latency, not HBM: 4096 channels x 4 sections = one warp per scheduler, "
                                    "y = (v + a1*y1) + a2*y2 is three dependent 5-cycle instructions per sample, so "
                                    "0.50 ms / 0.67 of the HBM roofline is the ceiling of EXACT arithmetic (DESIGN.md K1s, one group per SM; K1b)")
        del plan_o, xo, yo
        torch.cuda.empty_cache()
        peak_hbm = _peak_hbm()[0]
        try:
            # BASELINE configs[1] in FAST mode: too few channels for a lane each -> cut in time (warm-up form)
            Cs, Ts = WORKLOADS["c2"]
            plan_s, xs, ys, _, ms_s, launches_s, _ = device_rate("c2", coef="uniform", mode_name="fast")
            i_s = plan_s.info()
            K, L = i_s.warmup_samples, max(i_s.segment_samples, 1)
            ent = {"workload": workload_string("c2", args.layout) + ", FAST (FMA) mode, time-parallel",
                   "value": world * Cs * Ts / (ms_s * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_s,
                   "roofline_frac": BYTES_PER_SAMPLE * Cs * Ts / (ms_s * 1e-3) / 1e9 / peak_hbm,
                   "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE,
                   "moved_bytes_per_sample": (8 + 4.0 * K / L) if i_s.time_segments > 1 and K else (12 if i_s.time_segments > 1 else 8),
                   "kernel": i_s.kernel.decode(), "time_segments": i_s.time_segments, "segment_samples": i_s.segment_samples,
                   "warmup_samples": K, "launches_per_step": launches_s // max(args.steps, 1), "traffic": _traffic("c2_scan"),
                   "numerics": "FAST: FMA contraction + segments that start from a warmed-up state; held to the FAST bar of "
                               "tests/test_time_parallel.py (<= 3e-5 block-relative on this cascade, no further from float64 "
                               "than the reference)"}
            if rank == 0 and world == 1 and not args.no_cpu and args.layout == "planar":
                import flowz_oracle as fo                     # CPU leg: the checker, on four channels
                plan_s.reset()
                plan_s.process([xs], [ys])
                torch.cuda.synchronize()
                idx = [0, 1, Cs // 2, Cs - 1]
                xi = xs[idx].cpu().numpy()
                ref = fo.COracle(fo.biquad_cascade(SECTIONS), len(idx)).process([xi])[0]
                got = ys[idx].cpu().numpy().astype(np.float64)
                ent["parity_spot_check"] = {"max_block_rel_err_vs_oracle": float((np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)).max()),
                                            "channels_checked": idx}
            also["c2_scan"] = ent
            del plan_s, xs, ys
        except Exception as e:
            also["c2_scan"] = {"error": str(e)[:300]}
        torch.cuda.empty_cache()
        try:
            # the coefficients SURVEY.md 8(d) prescribes: every channel its own RBJ sections, f * (1 + c / C)
            Cn, Tn = WORKLOADS["ns"]
            plan_p, xp, yp, _, ms_p, _, _ = device_rate("ns", coef="per-channel", mode_name=args.mode)
            i_p = plan_p.info()
            also["ns_per_channel"] = {"workload": workload_string("ns", args.layout) + ", per-channel coefficients f*(1+c/C)",
                                      "value": world * Cn * Tn / (ms_p * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_p,
                                      "roofline_frac": BYTES_PER_SAMPLE * Cn * Tn / (ms_p * 1e-3) / 1e9 / peak_hbm,
                                      "kernel": i_p.kernel.decode(), "uniform_params": i_p.uniform_params, "mode": args.mode}
            del plan_p, xp, yp
        except Exception as e:
            also["ns_per_channel"] = {"error": str(e)[:300]}
        torch.cuda.empty_cache()
        try:
            # the same shape on the lane-per-channel kernel (K1; zg_plan_opts.section_warps = 1), what round 1 timed
            Cn, Tn = WORKLOADS["ns"]
            plan_k, xk, yk, _, ms_k, _, _ = device_rate("ns", mode_name=args.mode, section_warps=1)
            i_k = plan_k.info()
            also["ns_k1"] = {"workload": workload_string("ns", args.layout) + ", one lane per channel (section_warps = 1)",
                             "value": world * Cn * Tn / (ms_k * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_k,
                             "roofline_frac": BYTES_PER_SAMPLE * Cn * Tn / (ms_k * 1e-3) / 1e9 / peak_hbm,
                             "kernel": i_k.kernel.decode(), "mode": args.mode}
            del plan_k, xk, yk
        except Exception as e:
            also["ns_k1"] = {"error": str(e)[:300]}
        torch.cuda.empty_cache()
        also.update(other_configs(zg, wl, torch, dev, local, world, rank, args, barrier, dist))

    edge = None
    if world > 1 and not args.no_edge:
        try:
            edge = edge_step(zg, wl, torch, dist, dev, local, world, rank, args)
        except Exception as e:
            edge = {"error": str(e)[:300]}
        torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = _peak_hbm()
        achieved = BYTES_PER_SAMPLE * C * T / (ms_per_step * 1e-3) / 1e9       # per GPU, per launch
        line = {
            "metric": "Msamples/s, N-channel 4-section biquad chain", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, args.layout),
                       "coefficients": args.coef, "mode": args.mode, "parallelism": f"channel-shard x{world}",
                       "l2": f"inputs larger than L2 ({C * T * 4 / 2**20:.0f} MiB in + same out per GPU per step)",
                       "kernel": info.kernel.decode(), "lanes_per_channel": info.lanes_per_channel,
                       "threads_per_cta": info.threads_per_cta,
                       "stages": info.stages, "boxes": info.boxes, "smem_bytes": info.smem_bytes, "regs": info.regs_per_thread},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(args.workload), "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * C * T},
            "clocks": clk.summary(),
            "gpu_launches": launches,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if also is not None:
            line["also"] = also
        if edge is not None:
            line["edge"] = edge
        if world == 1 and not args.no_cpu:                  # CPU leg, part 2: the reference's loop on the host cores
            line["cpu_baseline"], _ = cpu_reference_rate()
            if parity is not None:
                line["cpu_baseline"]["parity_spot_check"] = parity
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns", choices=sorted(WORKLOADS))
    ap.add_argument("--coef", default="uniform", choices=["uniform", "per-channel"])
    ap.add_argument("--mode", default="exact", choices=["fast", "exact"])
    ap.add_argument("--layout", default="planar", choices=["planar", "interleaved"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the second workload")
    ap.add_argument("--no-edge", action="store_true", help="N > 1: skip the edge step (block resident on rank 0)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
