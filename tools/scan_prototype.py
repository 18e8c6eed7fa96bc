#!/usr/bin/env python
"""CPU prototype (numpy, CPU only) of a time-parallel evaluation of a LINEAR flowz graph for FEW, LONG channels -- the open
item behind BASELINE configs[1] (4096 channels x 65 536 samples leave most of a B200 idle with one lane per channel).

Two passes over S time segments per channel (S x the lanes):
  1. every segment ticks from ZERO state and keeps only its final state  z_s          (fully parallel)
  2. the true state at the start of segment s follows from  x_{s+1} = A^L x_s + z_s   (S steps per channel, tiny)
  3. every segment ticks again from its true initial state and writes its outputs      (fully parallel)
Pass 3 is the serial arithmetic of the reference started from a state that differs from the serial one by the rounding
of step 2, so the result is FAST-class (not bit-identical).  This script measures how far: block-relative error of the
segmented evaluation against the serial fp32 ticks and against float64, next to the reference's own distance to float64
(the noise floor of DESIGN.md 5).  Cost on the device: the input is read twice (12 B per sample instead of 8).

    python tools/scan_prototype.py [--segments 16] [--samples 65536] [--channels 64]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import flowz_oracle as fo          # noqa: E402  (tools may use the oracle: this is not product code)
import zignal_b200 as zg           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=16)
    ap.add_argument("--samples", type=int, default=65536)
    ap.add_argument("--channels", type=int, default=64)
    a = ap.parse_args()
    expr = fo.biquad_cascade(4)
    g = zg.compile(expr)
    A, B, Cm, D = g.state_space()
    C, T, S = a.channels, a.samples, a.segments
    L = T // S
    x = fo.noise(C, T, seed=3)
    serial = fo.COracle(expr, C).process([x])[0]
    truth = fo.biquad_cascade_f64(x, 4)

    # pass 1: zero-state final states (the oracle's state layout is not the product's: take them from the state-space
    # recursion in fp32, which is the same arithmetic up to association)
    def run_segment(x_seg, state0):
        """fp32 recursion state' = A state + B u for one segment of all channels; returns (y, final state)."""
        A32, B32, C32, D32 = (m.astype(np.float32) for m in (A, B, Cm, D))
        st = state0.astype(np.float32).copy()
        y = np.empty(x_seg.shape, np.float32)
        for t in range(x_seg.shape[1]):
            u = x_seg[:, t:t + 1]
            y[:, t] = (st @ C32.T + u @ D32.T)[:, 0]
            st = st @ A32.T + u @ B32.T
        return y, st

    zero = np.zeros((C, g.n_state), np.float32)
    finals = [run_segment(x[:, s * L:(s + 1) * L], zero)[1] for s in range(S)]
    AL = np.linalg.matrix_power(A, L).astype(np.float32)                      # step 2: A^L once per graph, in float64
    starts, cur = [], zero
    for s in range(S):
        starts.append(cur)
        cur = cur @ AL.T + finals[s]
    seg = np.concatenate([run_segment(x[:, s * L:(s + 1) * L], starts[s])[0] for s in range(S)], axis=1)
    one = run_segment(x, zero)[0]                                              # the same fp32 recursion, unsegmented

    def rel(y, ref):
        return float((np.abs(y.astype(np.float64) - ref).max(axis=1) / np.abs(ref).max(axis=1)).max())

    print(f"{C} channels x {T} samples, {S} segments of {L}; 4 x DF1 RBJ low-pass, block-relative error (max over channels)")
    print(f"  reference serial fp32 ticks   vs float64 : {rel(serial, truth):.3e}   (the noise floor)")
    print(f"  state-space fp32, one segment vs float64 : {rel(one, truth):.3e}")
    print(f"  state-space fp32, {S:3d} segments vs float64 : {rel(seg, truth):.3e}")
    print(f"  segmented vs unsegmented (cost of step 2) : {rel(seg, one.astype(np.float64)):.3e}")
    print(f"  segmented vs reference serial fp32 ticks  : {rel(seg, serial.astype(np.float64)):.3e}   (FAST is held to 3e-5)")


if __name__ == "__main__":
    main()
