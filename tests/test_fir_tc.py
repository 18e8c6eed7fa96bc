"""K3t: the dense FIR on tensor cores (tcgen05, 3xTF32; kernels/zg_fir_tc.cuh) -- FAST mode, planar, up to 256 taps.

Bar: 1e-5 block-relative against the oracle's fir_direct (the reference's left-to-right fp32 sum, flowz.hpp:769-772);
the kernel sums in blocked order with fp32 accumulators in TMEM and operands split into TF32 hi + lo, so it is not
bit-identical -- EXACT mode keeps the CUDA-core kernel, which is.  Measured on the benchmark workload: 3.1e-6 (the
CUDA-core FMA kernel: 2.6e-7); what dominates is the tensor core's fp32 accumulation, which truncates -- 36 accumulate
steps of the hi*hi chain per output (tools/fir_split_check.py; rounding the lo operands changes nothing)."""
import numpy as np
import pytest

import flowz_oracle as fo

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel_err(y, ref):
    den = np.abs(ref).max(axis=1)
    den = np.where(den == 0, 1.0, den)
    return (np.abs(y.astype(np.float64) - ref.astype(np.float64)).max(axis=1) / den).max()


def _plan(zg, h, C, **kw):
    return zg.compile(fo.fir_expr(h)).plan(channels=C, mode=zg.MODE_FAST, **kw)


@pytest.mark.parametrize("n_taps,C,T", [(256, 200, 1000), (256, 128, 256), (64, 70, 640), (100, 300, 2052), (2, 33, 512)])
def test_fir_tensor_core_parity(zg, n_taps, C, T):
    import torch
    h = fo.fir_taps(n_taps)
    x = fo.noise(C, T, seed=n_taps)
    plan = _plan(zg, h, C)
    y = plan.process([zg.to_block(x)])[0]
    torch.cuda.synchronize()
    assert plan.info().kernel.decode() == f"zg_fir_tc<{n_taps} taps,3xtf32,planar>" and plan.info().jit == 0
    err = _rel_err(y.cpu().numpy(), fo.fir_direct(x, h))
    assert err <= TOL, err


def test_fir_tensor_core_streams_like_ticks(zg):
    """Blocks continue the stream: the delay line (oldest first, rotate_push_back flowz.hpp:130-148) is written by the
    state kernel after every block and read back for the history of the next, also when a block is shorter than
    the line."""
    import torch
    C, T = 96, 1536
    h = fo.fir_taps(256)
    x = fo.noise(C, T, seed=77)
    ref = fo.fir_direct(x, h)
    plan = _plan(zg, h, C)
    outs = []
    t0 = 0
    for n in (512, 256, 768):                             # 256 < 255 + 1 taps of history: part of the old line survives
        outs.append(plan.process([zg.to_block(x[:, t0:t0 + n])])[0].cpu().numpy())
        t0 += n
    torch.cuda.synchronize()
    assert _rel_err(np.concatenate(outs, axis=1), ref) <= TOL
    assert np.array_equal(plan.get_state(), x[:, -255:].T)
    # short blocks fall back to the CUDA-core FMA kernel and share the same delay line
    y_short = plan.process([zg.to_block(fo.noise(C, 100, seed=78))])[0].cpu().numpy()
    want = fo.fir_direct(fo.noise(C, 100, seed=78), h, history=x[:, -255:])
    assert _rel_err(y_short, want) <= TOL
    plan.reset()
    assert _rel_err(plan.process([zg.to_block(x)])[0].cpu().numpy(), ref) <= TOL


def test_fir_tensor_core_unit_taps_are_exact(zg):
    """Unit taps on a small-integer ramp: every operand is a TF32 number, every partial sum an exact fp32 integer --
    delay indexing through the Toeplitz windows must be exact."""
    n, C, T = 200, 130, 1024
    x = np.tile((np.arange(T) % 97 + 1).astype(np.float32), (C, 1))
    x[1] = x[1][::-1]
    y = _plan(zg, np.ones(n, np.float32), C).process([zg.to_block(x)])[0].cpu().numpy()
    for c in (0, 1, 129):
        cs = np.concatenate([np.zeros(1), np.cumsum(x[c].astype(np.float64))])
        t = np.arange(T)
        assert np.array_equal(y[c], (cs[t + 1] - cs[np.maximum(t + 1 - n, 0)]).astype(np.float32))


def test_fir_tensor_core_through_process_host(zg):
    h = fo.fir_taps(256)
    x = fo.noise(4096, 2048, seed=9)
    plan = _plan(zg, h, 4096)
    y = plan.process_host([x])[0]
    assert b"zg_fir_tc" in plan.info().kernel
    idx = [0, 127, 128, 2049, 4095]
    assert _rel_err(y[idx], fo.fir_direct(x[idx], h)) <= TOL


def test_full_size_config4_fir256_on_tensor_cores(zg):
    """BASELINE configs[3] at its full shape (32 768 channels x 8192 samples x 256 taps) in FAST mode: sampled
    channels against the oracle, every channel against the bit-identical EXACT kernel, identical inputs give
    identical channels, linearity."""
    import torch
    C, T = 32768, 8192
    h = fo.fir_taps(256)
    g = zg.compile(fo.fir_expr(h))
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((C, T), generator=gen, device="cuda") * 2 - 1
    x[1::2] = x[0::2]
    plan = g.plan(channels=C, mode=zg.MODE_FAST)
    y = plan.process([x])[0]
    torch.cuda.synchronize()
    assert b"zg_fir_tc" in plan.info().kernel
    assert torch.equal(y[0::2], y[1::2])
    idx = [0, 1, 127, 128, 4097, 32766, 32767]
    assert _rel_err(y[idx].cpu().numpy(), fo.fir_direct(x[idx].cpu().numpy(), h)) <= TOL
    ye = g.plan(channels=C, mode=zg.MODE_EXACT).process([x])[0]
    den = ye.abs().amax(dim=1)
    assert float(((y - ye).abs().amax(dim=1) / den).max()) <= TOL
    y2 = g.plan(channels=C, mode=zg.MODE_FAST).process([x * 0.5])[0]
    assert float(((y2 - 0.5 * y).abs().amax(dim=1) / den).max()) <= 1e-6


def test_fir_tensor_cores_can_be_declined(zg):
    """zg_plan_opts.fir_tensor_cores = 1: FAST mode on the CUDA-core FMA kernel (an order of magnitude closer to the
    reference's sum, a third of the speed)."""
    h = fo.fir_taps(256)
    x = fo.noise(64, 1024, seed=12)
    plan = _plan(zg, h, 64, fir_tensor_cores=1)
    y = plan.process([zg.to_block(x)])[0].cpu().numpy()
    assert plan.info().kernel.decode() == "zg_fir<256 taps,fma,planar>"
    assert _rel_err(y, fo.fir_direct(x, h)) <= 1e-6
