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


# ---- graphs beyond the reference on the device --------------------------------------------------------------------
# Feedbacks the reference cannot split (nested loops, parallel combiners inside a loop: TODO.md:11-29, the disabled
# expectation test/tests.cpp:59) and zero-input sources (TODO.md:65) lower to the same kind of tick program; here the
# kernels generated from them run on the GPU.  The oracle has no opinion on these graphs, so the checker is the netlist
# evaluator (tests/netlist_flowz.py), ticked for a few of the channels.

FIXED_BEYOND = [
    "~~( _1 + _2 + 1.0f |= _1[_1] )",                          # nested feedback, no external input: 0, 1, 3, 7, ...
    "~( (0.5f*_1 + _2) |= ~(0.25f*_1 + _2 |= _1[_1]) )",       # a loop inside a loop, one input
    "~((_1[_2] |= 0.5f*_2) | (_1 - _2))",                      # parallel combiner inside a loop
    "~(0.5f*_1[_1] + 1.0f)",                                   # zero-input source (make_front<0>): a constant-driven one-pole
    "~(0x1.fcp0f*_1[_1] - _1[_2] + 0.125f)",                   # zero-input recursive oscillator with an offset
    "~(-0.5f*_1[_2] + 0.25f) |= (_1 , _1[_1])",                # zero inputs, two outputs
]


def _beyond_graph(zg, rng):
    """A random all-fp32 graph the product compiles and the oracle (= the reference's rules) refuses."""
    import netlist_flowz as nl
    for _ in range(20000):
        expr = _gen(rng, rng.randint(3, 5), rng.randint(1, 2), consts=["0.5f", "0.25f", "-0.75f", "0x1p-1f", "-0.5f"], ops="+-*")
        if "~" not in expr:
            continue
        try:
            g = zg.compile(expr)
        except zg.ZgError:
            continue
        if not g.all_f32 or g.n_out < 1 or g.n_state > 48:
            continue
        try:
            res = fo.Oracle(expr).tick(*([0.0] * g.n_in))
            if not any(r is fo.BOTTOM for r in res):
                continue                                 # the reference compiles it: covered by the test above
        except Exception:
            pass
        try:
            nl.Netlist(expr).tick(*([0.0] * g.n_in))
        except Exception:
            continue
        return expr, g
    raise AssertionError("generator produced no graph beyond the reference")


def _run_any(zg, g, x, T, layout, blocks, C):
    """Like test_gpu_parity._run, for graphs with or without inputs."""
    import torch
    import zignal_b200
    plan = g.plan(channels=C, mode=zg.MODE_EXACT, layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED)
    outs = [[] for _ in range(g.n_out)]
    t0 = 0
    for n in blocks:
        ins = [zignal_b200.to_block(xk[:, t0:t0 + n].T if layout == "interleaved" else xk[:, t0:t0 + n]) for xk in x]
        ys = plan.process(ins, n_samples=n)
        torch.cuda.synchronize()
        for o, y in zip(outs, ys):
            y = y.cpu().numpy()
            o.append(y.T if layout == "interleaved" else y)
        t0 += n
    return [np.concatenate(o, axis=1) for o in outs], plan


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(2))
def test_graphs_beyond_the_reference_on_the_device_equal_the_netlist(zg, seed):
    import netlist_flowz as nl
    rng = random.Random(9100 + seed)
    C, T = 45, 72
    exprs = [(e, zg.compile(e)) for e in FIXED_BEYOND[seed::2]] + [_beyond_graph(zg, rng) for _ in range(4)]
    ran = zero_input = 0
    for i, (expr, g) in enumerate(exprs):
        x = [fo.noise(C, T, seed=300 * seed + 10 * i + k) for k in range(g.n_in)]
        layout = "planar" if (seed + i) % 2 == 0 else "interleaved"
        try:
            ys, plan = _run_any(zg, g, x, T, layout, [29, T - 29], C)
        except zg.ZgError as e:
            if e.status == zg.ZG_ERR_UNSUPPORTED:
                continue
            raise AssertionError(f"{expr}: {e}")
        assert plan.info().jit == 1
        for c in (0, 17, C - 1):
            net = nl.Netlist(expr)
            ticks = [net.tick(*[float(xk[c, t]) for xk in x]) for t in range(T)]
            for o in range(g.n_out):
                w = np.array([tk[o][1] for tk in ticks], np.float32)
                y = ys[o][c]
                same = (y.view(np.uint32) == w.view(np.uint32)) | (np.isnan(y) & np.isnan(w))
                assert same.all(), f"{expr} [{layout}] channel {c} output {o}: {np.argwhere(~same)[:4].ravel().tolist()}"
        ran += 1
        zero_input += g.n_in == 0
    assert ran >= 5 and zero_input >= 1
