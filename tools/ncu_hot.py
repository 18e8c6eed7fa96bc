#!/usr/bin/env python
"""Per-instruction view of an `ncu --set full --import-source on` report: SASS lines with their stall
samples and execution counts, hottest regions first.  Runs here, no GPU.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [--top 60] [--range A:B]   (line indices of the SASS listing)"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    rng = sys.argv[sys.argv.index("--range") + 1] if "--range" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {k: hdr.index(k) for k in hdr}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    body = rows[2:]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    print(f"# {rows[0][1]}   total samples {tot}")
    def fmt(i, r):
        s = int(r[ix["# Samples"]] or 0)
        st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)
        st = " ".join(f"{n}:{v}" for v, n in st[:3] if v)
        return f"{i:5d} {s:6d} {100*s/tot:5.1f}% x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip():60s} {st}"
    if rng:
        a, b = (int(v) for v in rng.split(":"))
        for i in range(a, b): print(fmt(i, body[i]))
        return
    if top:
        order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
        for i in sorted(order): print(fmt(i, body[i]))
        return
    # default: cumulative samples per 64-instruction window
    W = 64
    for a in range(0, len(body), W):
        s = sum(int(r[ix["# Samples"]] or 0) for r in body[a:a+W])
        ex = max(int(r[ix["Instructions Executed"]] or 0) for r in body[a:a+W])
        print(f"lines {a:5d}-{a+W-1:5d}  samples {s:7d} {100*s/tot:5.1f}%   max exec {ex}")

if __name__ == "__main__":
    main()
