"""Random fp32 flowz graphs on the device (generated kernels, EXACT mode) against the oracle's C tick, bit for bit.

The graphs come from the same grammar-driven generator as tests/test_fuzz_frontend.py with float constants only, so
every one of them is a valid all-fp32 graph for the block evaluator: series, parallel, fan-out, (nested) feedback, delays
up to 3 on any wire.  Each is run for 70 ragged channels x 80 samples in two uneven blocks (state carries over), planar
and interleaved alternating.  Feedback graphs may grow without bound; where the oracle's value is a NaN the device's
must be one too (x86 and the GPU differ in the NaN they produce), everything else must match in every bit."""
import random

import numpy as np
import pytest

import flowz_oracle as fo
from test_fuzz_frontend import _gen
from test_gpu_parity import _run

GRAPHS_PER_SEED = 6


def _valid_f32_graph(zg, rng, want_feedback):
    for _ in range(5000):
        expr = _gen(rng, rng.randint(3, 5), rng.randint(1, 3), consts=["0.5f", "0.25f", "-0.75f", "0x1p-1f", "1.5f", "-1.0f"], ops="+-*")
        if ("~" in expr) != want_feedback or (not want_feedback and "|" not in expr and "," not in expr):
            continue                                 # half the graphs recurse, the others at least route wires
        try:
            g = zg.compile(expr)
        except zg.ZgError:
            continue
        if not g.all_f32 or g.n_in < 1 or g.n_out < 1:
            continue
        try:
            o = fo.Oracle(expr)
            if any(r is fo.BOTTOM for r in o.tick(*([0.0] * g.n_in))):
                continue
        except Exception:
            continue
        return expr, g
    raise AssertionError("generator produced no valid graph")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_random_graphs_on_the_device_equal_the_oracle(zg, seed):
    rng = random.Random(7000 + seed)
    C, T = 70, 80
    ran = 0
    for i in range(GRAPHS_PER_SEED):
        expr, g = _valid_f32_graph(zg, rng, want_feedback=i % 2 == 0)
        x = [fo.noise(C, T, seed=100 * seed + 10 * i + k) for k in range(g.n_in)]
        layout = "planar" if (seed + i) % 2 == 0 else "interleaved"
        try:
            ys, plan = _run(zg, expr, x, zg.MODE_EXACT, layout=layout, blocks=[37, T - 37])
        except zg.ZgError as e:
            if e.status == zg.ZG_ERR_UNSUPPORTED:        # a documented limit of the device path (state, wires)
                continue
            raise AssertionError(f"{expr}: {e}")
        want = fo.COracle(expr, C).process(x)
        assert len(ys) == len(want), expr
        for y, w in zip(ys, want):
            same = y.view(np.uint32) == w.view(np.uint32)
            both_nan = np.isnan(y) & np.isnan(w)
            assert (same | both_nan).all(), f"{expr} [{layout}]: {np.argwhere(~(same | both_nan))[:4].tolist()}"
        ran += 1
    assert ran >= GRAPHS_PER_SEED // 2
