#!/usr/bin/env python
"""In-order single-warp issue model of a SASS loop: one instruction per cycle, an instruction waits until its
source registers are ready (fixed latency 4 for FP32 / integer ALU, 30 for LDS).  Estimates the cycles of the
hot loop of a kernel that runs one warp per scheduler (K1b), i.e. how well ptxas interleaved the recurrence.
    python tools/sass_sim.py file.sass   (lines: /*addr*/ OP operands ;)   prints the loop found and its cycles"""
import re, sys
lines = [l.rstrip() for l in open(sys.argv[1])]
ins = []
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_ix = {a: i for i, (a, _) in enumerate(ins)}
# hottest loop = backward branch whose body holds the most STS.128 and no other backward branch target inside
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
    if m and "BRA.DIV" not in t:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_ix:
            body = ins[addr_ix[tgt]:i + 1]
            n = sum("STS.128" in x for _, x in body)
            if n >= 8 and (best is None or len(body) < len(best)): best = body
if best is None: sys.exit("no loop with 8 STS.128 found")
def regs(tok):
    out = []
    for r in re.findall(r"\bR(\d+)\b", tok):
        out.append(int(r))
    return out
FPLAT = int(sys.argv[2]) if len(sys.argv) > 2 else 4
LAT = {"LDS": 30}
ready = {}
t = 0
stalls = 0
for _ in range(3):                      # iterate the loop a few times to reach steady state
    start = t
    for a, txt in best:
        op = txt.split()[0]
        if op.startswith("@"): op = txt.split()[1]
        body = txt[txt.index(op) + len(op):]
        parts = [p.strip() for p in body.split(",")]
        dst = regs(parts[0]) if parts and not op.startswith(("STS", "BRA", "NOP", "BSSY", "BSYNC", "WARPSYNC")) else []
        srcs = [r for p in (parts[1:] if dst else parts) for r in regs(p)]
        wide = 4 if ".128" in op else 2 if ".64" in op else 1
        if op.startswith("STS"): srcs = srcs + [r + k for r in regs(parts[-1]) for k in range(1, wide)]
        need = max([ready.get(r, 0) for r in srcs] + [t])
        stalls += need - t
        t = need + 1
        lat = LAT.get(op.split(".")[0], FPLAT)
        for d in dst:
            for k in range(wide if op.startswith("LDS") else 1): ready[d + k] = t - 1 + lat
    last = t - start
print(f"loop of {len(best)} instructions (8 iterations): {last} cycles per pass in steady state = {last/8:.1f} per iteration")
