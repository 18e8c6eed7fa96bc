#!/usr/bin/env python
"""Small invocations of every device kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize/device_check.py
K1, K1s, K3, generated kernels, bf16 storage, a warm-up time-segmented launch, a comb + echo graph whose delay lines live in
HBM as rings, an in-place block, the two-pass time-segmented form and the tensor-core FIR; `--k1b` runs ONLY the
section-parallel biquad kernel (whose intra-warp hand-over through shared memory racecheck reports as warp-level
warnings -- see profiles/README.md), `--smoke` runs __graft_entry__.smoke() first."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import __graft_entry__ as entry
import flowz_oracle as fo
import zignal_b200 as zg

# The race checker follows a thread's OWN mbarrier arrivals only, so for it every lane of a K1s warp arrives itself
# (ZG_TUNE_SPLIT_ARRIVE=1, kernels/zg_biquad_split.cuh kAllArrive); the production form -- lane 0 arrives for the warp
# after a __syncwarp -- is what memcheck / synccheck see (ZG_SANITIZER=memcheck|synccheck), and what `--k1s-elected`
# forces under racecheck too (it then reports the hand-over of every box as a hazard).
if "--k1s-elected" not in sys.argv and os.environ.get("ZG_SANITIZER", "racecheck") == "racecheck":
    os.environ["ZG_TUNE_SPLIT_ARRIVE"] = "1"

C, T = 64, 2048
x = fo.noise(C, T, seed=3)
xd = zg.to_block(x)
expr = fo.biquad_cascade(4)
ref = fo.COracle(expr, C).process([x])[0]
if "--smoke" in sys.argv:
    entry.smoke()
if "--k1b" in sys.argv:
    lanes = zg.compile(expr).plan(channels=C, lanes_per_channel=4)
    assert np.array_equal(lanes.process([xd])[0].cpu().numpy(), ref) and lanes.info().lanes_per_channel == 4
    print("device_check ok (K1b only)")
    sys.exit(0)

# K1 (EXACT, product-reusing tick; FAST; FAST cut in time with warm-up), K3, generated kernel, bf16 storage
k1 = zg.compile(expr).plan(channels=C, lanes_per_channel=1)
assert np.array_equal(k1.process([xd])[0].cpu().numpy(), ref), "K1"
x8 = fo.noise(C, 8192, seed=5)
cut = zg.compile(expr).plan(channels=C, mode=zg.MODE_FAST, time_parallel=zg.TP_WARMUP)
y8 = cut.process([zg.to_block(x8)])[0].cpu().numpy()
r8 = fo.COracle(expr, C).process([x8])[0]
assert cut.info().time_segments >= 2 and (np.abs(y8 - r8).max(axis=1) / np.abs(r8).max(axis=1)).max() <= 3e-5, "warm-up segments"
osc = "~(0x1.fcp0f*_1[_1] - _1[_2] + _2) >> ~(_2 + 0.9f*_1[-1])"
d = np.zeros((C, T), np.float32); d[:, 0] = 1
yo = zg.compile(osc).plan(channels=C, input_kind=[zg.IN_DIRAC]).process([None], n_samples=T)[0]
assert np.array_equal(yo.cpu().numpy(), fo.COracle(osc, C).process([d])[0]), "generated kernel"
taps = fo.fir_taps(64)
assert np.array_equal(zg.compile(fo.fir_expr(taps)).plan(channels=C).process([xd])[0].cpu().numpy(), fo.fir_direct(x, taps)), "K3"
yb = zg.compile(expr).plan(channels=C, io_dtype=zg.BF16).process([xd.to(torch.bfloat16)])[0]
refb = fo.COracle(expr, C).process([fo.bf16_round(x)])[0]
assert np.array_equal(yb.view(torch.int16).cpu().numpy().view(np.uint16), fo.bf16_bits(refb)), "bf16 storage"

# K1s: sections spread over the warps of a group (box hand-over through mbarriers; rows cut between groups)
os.environ["ZG_TUNE_SPLIT_G"] = "3"          # three groups per CTA, as on many channels (328 channels alone would run one group per SM)
ks = zg.compile(expr).plan(channels=328, lanes_per_channel=1, section_warps=2)
xs = fo.noise(328, 1504, seed=6)
assert np.array_equal(ks.process([zg.to_block(xs)])[0].cpu().numpy(), fo.COracle(expr, 328).process([xs])[0]), "K1s"
assert b"zg_biquad_df1_split<4,exact,planar,4 warps per group>" in ks.info().kernel
os.environ.pop("ZG_TUNE_SPLIT_G")
# ... and its few-channel form (one group per SM, four boxes per hand-over), what an auto plan of 64 channels runs on a long block
kf = zg.compile(expr).plan(channels=C)
xf8 = fo.noise(C, 8192, seed=9)
assert np.array_equal(kf.process([zg.to_block(xf8)])[0].cpu().numpy(), fo.COracle(expr, C).process([xf8])[0]), "K1s, few channels"
assert b"boxes per hand-over" in kf.info().kernel

# long delay lines: rings in HBM
comb = "~(_2 + 0.5f*_1[_441]) |= (_1 + 0.25f*_1[_1000])"
y = zg.compile(comb).plan(channels=C).process([xd])[0].cpu().numpy()
assert np.array_equal(y, fo.COracle(comb, C).process([x])[0]), "ring graph"

# in place
buf = zg.to_block(x)
zg.compile(expr).plan(channels=C, lanes_per_channel=1).process([buf], outputs=[buf])
torch.cuda.synchronize()
assert np.array_equal(buf.cpu().numpy(), fo.COracle(expr, C).process([x])[0]), "in place"

# two-pass time segments (pass 1, boundary fix-up, pass 2)
p2 = zg.compile(expr).plan(channels=C, mode=zg.MODE_FAST, time_parallel=zg.TP_TWO_PASS)
y2 = p2.process([xd])[0].cpu().numpy()
assert p2.info().time_segments >= 2 and (np.abs(y2 - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() <= 3e-5, "two-pass"

# tensor-core FIR
h = fo.fir_taps(256)
pf = zg.compile(fo.fir_expr(h)).plan(channels=200, mode=zg.MODE_FAST)
xf = fo.noise(200, 1024, seed=4)
yf = pf.process([zg.to_block(xf)])[0].cpu().numpy()
rf = fo.fir_direct(xf, h)
assert b"zg_fir_tc" in pf.info().kernel and (np.abs(yf - rf).max(axis=1) / np.abs(rf).max(axis=1)).max() <= 1e-5, "fir tc"
print("device_check ok")
