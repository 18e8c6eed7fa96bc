#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary committed under profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt
(runs here, no GPU needed: `ncu -i ... --page raw --csv`)"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d: print(f"  {k:75s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        print("  top warp stall reasons (warps stalled per issue-active cycle):")
        for v, k in stalls[:6]:
            print(f"    {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:.3f}")
        rd, wr = float(d.get("dram__bytes_read.sum", 0)), float(d.get("dram__bytes_write.sum", 0))
        t = float(d.get("gpu__time_duration.sum", 0))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        b = rd * scale.get(u.get("dram__bytes_read.sum"), 1) + wr * scale.get(u.get("dram__bytes_write.sum"), 1)
        tt = t * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u.get("gpu__time_duration.sum"), 1)
        if tt: print(f"  => dram traffic {b/1e9:.4f} GB per launch, {b/tt/1e9:.0f} GB/s under the profiler (serialised, cold cache)")

if __name__ == "__main__":
    main()
