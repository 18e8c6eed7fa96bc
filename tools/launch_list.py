#!/usr/bin/env python
"""ncu launch-list CSV (gpu__time_duration.sum per launch) -> compact table for profiles/.
    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Value", "Metric Unit")}
tot = {}
print(f"{'id':>4} {'time':>12} unit  grid/block  kernel")
for r in rows[1:]:
    name = r[ix["Kernel Name"]]
    short = name if len(name) < 90 else name[:87] + "..."
    t = float(r[ix["Metric Value"]].replace(",", ""))
    tot[short] = tot.get(short, 0) + t
    print(f"{r[ix['ID']]:>4} {t:12.0f} {r[ix['Metric Unit']]:4s} {r[ix['Grid Size']]}/{r[ix['Block Size']]}  {short}")
s = sum(tot.values())
print("\nshare of listed device time (cold cache, serialised under the profiler):")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {100 * v / s:6.2f} %  {k}")
