#!/usr/bin/env python
"""How does tcgen05.mma.kind::tf32 turn an fp32 operand into TF32?  (GPU box only.)  ZG_TUNE_FIR_SPLIT selects what the
split warps of kernels/zg_fir_tc.cuh leave in shared memory: 0 = hi rewritten as cvt.rna.tf32(x) (the default: exact
whatever the hardware does), 1 = hi left as the raw fp32 block with lo = x - trunc(x), 2 = hi raw with lo = x - rna(x).
Measured on B200: 1 agrees with 0 (3e-6), 2 is off by 7e-4 -- the tensor core TRUNCATES.  Also prints the error of the
CUDA-core FMA kernel (K3, FAST) on the same block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import zignal_b200 as zg
import flowz_oracle as fo

h = fo.fir_taps(256); x = fo.noise(256, 4096, seed=1); ref = fo.fir_direct(x, h)
err = lambda y: float((np.abs(y - ref).max(axis=1) / np.abs(ref).max(axis=1)).max())
for m in ("0", "1", "2"):
    os.environ["ZG_TUNE_FIR_SPLIT"] = m
    y = zg.compile(fo.fir_expr(h)).plan(channels=256, mode=zg.MODE_FAST).process([zg.to_block(x)])[0].cpu().numpy()
    print("split mode", m, "block-relative error vs the oracle", err(y))
os.environ["ZG_TUNE_FIR_SPLIT"] = "0"; os.environ["ZG_TUNE_FIR_TC"] = "1"
y = zg.compile(fo.fir_expr(h)).plan(channels=256, mode=zg.MODE_FAST).process([zg.to_block(x)])[0].cpu().numpy()
print("CUDA-core FMA kernel (K3 FAST)", err(y))
