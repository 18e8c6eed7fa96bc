"""K1s (kernels/zg_biquad_split.cuh): the sections of a biquad cascade spread over the warps of a group that share one
ring of tiles.  The arithmetic of a section is the lane-per-channel kernel's (BiquadDf1Cascade::tick), so everything here
is BIT-IDENTICAL: to the oracle in EXACT mode, and to K1 in either mode.  Through the C ABI on cuda:0."""
import numpy as np
import pytest

import flowz_oracle as fo

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _run(zg, expr, x, mode, section_warps, params=None, blocks=None, in_place=False):
    torch = _torch()
    import zignal_b200
    C, T = x.shape
    plan = zg.compile(expr).plan(channels=C, mode=mode, lanes_per_channel=1, section_warps=section_warps)
    for i, p in enumerate(params or []):
        plan.set_param(i, p)
    outs, t0 = [], 0
    for n in (blocks or [T]):
        xb = zignal_b200.to_block(x[:, t0:t0 + n])
        y = plan.process([xb], n_samples=n, outputs=[xb] if in_place else None)[0]
        torch.cuda.synchronize()
        outs.append(y.cpu().numpy())
        t0 += n
    return np.concatenate(outs, axis=1), plan


def _is_split(plan):
    return b"zg_biquad_df1_split" in plan.info().kernel


@pytest.mark.parametrize("sections", [2, 3, 4, 6, 8])
def test_split_exact_is_bit_identical_to_the_oracle(zg, sections):
    C, T = 200, 1504                    # ragged channel count; 47 boxes: three tiles of 14 and one of 5
    x = fo.noise(C, T, seed=sections)
    expr = fo.biquad_cascade(sections)
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2)
    assert _is_split(plan), plan.info().kernel
    assert plan.info().launches == 1
    ref = fo.COracle(expr, C).process([x])[0]
    assert np.array_equal(y, ref)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_split_equals_lane_per_channel_kernel_many_rows_per_group(zg, mode):
    """more channel groups than persistent groups: every group walks several rows, the tile stream runs through the row
    boundaries (state written and read in between), the last round is partly empty"""
    torch = _torch()
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    C = 32 * (2 * sm * 3 + 37) + 5
    T = 928                             # 29 boxes: ragged last tile in every row
    m = zg.MODE_EXACT if mode == "exact" else zg.MODE_FAST
    x = fo.noise(C, T, seed=3)
    expr = fo.biquad_cascade(4)
    y, plan = _run(zg, expr, x, m, 2)
    assert _is_split(plan)
    z, plan1 = _run(zg, expr, x, m, 1)
    assert not _is_split(plan1)
    assert np.array_equal(y, z)
    assert np.array_equal(np.asarray(plan.get_state()), np.asarray(plan1.get_state()))


def test_split_streams_like_ticks_and_shares_state_rows_with_k1(zg):
    """consecutive blocks continue the stream; the state a split launch leaves is the state K1 continues from"""
    torch = _torch()
    import zignal_b200
    C, T = 96, 4096
    x = fo.noise(C, T, seed=11)
    expr = fo.biquad_cascade(4)
    ref = fo.COracle(expr, C).process([x])[0]
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2, blocks=[1024, 2048, 1024])
    assert _is_split(plan)
    assert np.array_equal(y, ref)
    # first half on K1s, state moved into a K1 plan, second half there
    a = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=1, section_warps=2)
    b = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=1, section_warps=1)
    y0 = a.process([zignal_b200.to_block(x[:, :2048])], n_samples=2048)[0].cpu().numpy()
    b.set_state(a.get_state())
    y1 = b.process([zignal_b200.to_block(x[:, 2048:])], n_samples=2048)[0].cpu().numpy()
    assert _is_split(a) and not _is_split(b)
    assert np.array_equal(np.concatenate([y0, y1], axis=1), ref)


@pytest.mark.parametrize("C", [160, 800])
def test_split_per_channel_coefficients(zg, C):
    """$k parameters, one value per channel (b0 == b2 per channel: the product-reusing tick), and a set that is not
    symmetric (the plain tick); 160 channels = one group per SM, 800 = three per CTA"""
    T = 1024
    x = fo.noise(C, T, seed=5)
    expr = fo.biquad_cascade_params(4)
    for sym in (True, False):
        rng = np.random.default_rng(17)
        params = []
        for k in range(4):
            f = 200.0 * (k + 1) * (1.0 + np.arange(C) / C)
            b0, b1, b2, a1, a2 = fo.rbj_lowpass(f, 0.7 + 0.1 * k, 48000.0)
            if not sym:
                b2 = (b2 * (1.0 + 0.01 * rng.standard_normal(C))).astype(np.float32)
            params += [np.asarray(v, np.float32) for v in (b0, b1, b2, a1, a2)]
        y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2, params=params)
        assert _is_split(plan)
        # (160 channels run one group per SM: that form evaluates plain sections, it has no use for the product reuse)
        assert b"boxes per hand-over" in plan.info().kernel or (b"+b0=b2" in plan.info().kernel) == sym
        prm = np.stack([np.broadcast_to(p, (C,)) for p in params], axis=1)
        ref = fo.COracle(expr, C, params=prm).process([x])[0]
        assert np.array_equal(y, ref)


@pytest.mark.parametrize("boxes", [1, 3, 5])
def test_split_rows_cut_between_groups_are_bit_identical(zg, boxes, monkeypatch):
    """every group gets the same number of tiles, so rows are cut between groups: the head piece of a row leaves its
    delay lines for the group that runs the tail (EXACT: the same ticks in the same order, still bit-identical).  Small
    tiles move the cuts around; two blocks per plan: the flags and the CTA tickets run on from launch to launch"""
    monkeypatch.setenv("ZG_TUNE_BOXES", str(boxes))
    monkeypatch.setenv("ZG_TUNE_SPLIT_G", "3")      # (so few channels would run one group per SM, whole rows each)
    C, T = 328, 1504                    # 11 channel groups over 9 groups of warps; 47 boxes per row
    x = fo.noise(C, T, seed=boxes)
    expr = fo.biquad_cascade(4)
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2, blocks=[736, 768])
    assert _is_split(plan), plan.info().kernel
    assert plan.info().boxes == boxes
    ref = fo.COracle(expr, C).process([x])[0]
    assert np.array_equal(y, ref)


def test_split_many_launches_of_one_plan(zg):
    torch = _torch()
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    C = 32 * (7 * sm + 11)
    T = 640
    x = fo.noise(C, 4 * T, seed=21)
    expr = fo.biquad_cascade(4)
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2, blocks=[T, T, T, T])
    assert _is_split(plan)
    z, plan1 = _run(zg, expr, x, zg.MODE_EXACT, 1)
    assert not _is_split(plan1)
    assert np.array_equal(y, z)
    assert np.array_equal(np.asarray(plan.get_state()), np.asarray(plan1.get_state()))


def test_split_form_for_the_race_checker_is_the_same_arithmetic(zg, monkeypatch):
    """ZG_TUNE_SPLIT_ARRIVE=1: every lane arrives on the hand-over barriers itself (what compute-sanitizer racecheck runs,
    tools/sanitize/device_check.py) -- same results"""
    monkeypatch.setenv("ZG_TUNE_SPLIT_ARRIVE", "1")
    C, T = 328, 1504
    x = fo.noise(C, T, seed=31)
    expr = fo.biquad_cascade(4)
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2)
    assert _is_split(plan)
    assert np.array_equal(y, fo.COracle(expr, C).process([x])[0])


def test_split_replayed_from_a_cuda_graph(zg):
    """nothing about a K1s launch lives on the host (the CTA tickets and the epoch of the row flags are kept by the kernel
    itself), so captured launches can be replayed: every replay continues the stream"""
    torch = _torch()
    C, T, NBLK = 1000, 256, 4
    expr = fo.biquad_cascade(4)
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=1, section_warps=2)
    x = fo.noise(C, 3 * NBLK * T, seed=77)
    xin = torch.empty((C, NBLK * T), device="cuda")
    yout = torch.empty_like(xin)
    plan.process([xin[:, :T]], [yout[:, :T]])                 # warm: buffers allocated, kernel attributes set
    assert _is_split(plan)
    plan.reset()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for b in range(NBLK):
                plan.process([xin[:, b * T:(b + 1) * T]], [yout[:, b * T:(b + 1) * T]])
    plan.reset()
    outs = []
    for rep in range(3):
        xin.copy_(torch.from_numpy(x[:, rep * NBLK * T:(rep + 1) * NBLK * T]))
        graph.replay()
        torch.cuda.synchronize()
        outs.append(yout.cpu().numpy().copy())
    idx = [0, 31, 32, 517, 999]
    ref = fo.COracle(expr, len(idx)).process([x[idx]])[0]
    assert np.array_equal(np.concatenate(outs, axis=1)[idx], ref)


def test_split_few_channels_one_group_per_sm(zg):
    """few channels, EXACT, long blocks (BASELINE configs[1] in small): the auto rule runs a warp per section with one
    group per SM, whole rows per group, four boxes per hand-over -- bit-identical to the oracle, streams like ticks
    (second block ragged: 131 boxes), shares its state rows with the lanes kernel K1b"""
    torch = _torch()
    import zignal_b200
    C = 100
    T1, T2 = 8192, 4192
    x = fo.noise(C, T1 + T2, seed=41)
    expr = fo.biquad_cascade(4)
    ref = fo.COracle(expr, C).process([x])[0]
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT)
    assert plan.info().lanes_per_channel == 4                 # sized for K1b ...
    y1 = plan.process([zignal_b200.to_block(x[:, :T1])], n_samples=T1)[0].cpu().numpy()
    i = plan.info()
    assert b"zg_biquad_df1_split<4,exact,planar,4 warps per group,4 boxes per hand-over>" in i.kernel and i.lanes_per_channel == 1
    y2 = plan.process([zignal_b200.to_block(x[:, T1:])], n_samples=T2)[0].cpu().numpy()
    assert np.array_equal(np.concatenate([y1, y2], axis=1), ref)
    # a short block of the same plan runs on K1b; the state rows are the same
    k1b = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=4)
    k1b.process([zignal_b200.to_block(x[:, :T1])], n_samples=T1)
    plan2 = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT)
    plan2.set_state(k1b.get_state())
    z2 = plan2.process([zignal_b200.to_block(x[:, T1:])], n_samples=T2)[0].cpu().numpy()
    assert np.array_equal(z2, ref[:, T1:])
    # per-channel coefficients
    pexpr = fo.biquad_cascade_params(4)
    params = []
    for k in range(4):
        f = 300.0 * (k + 1) * (1.0 + np.arange(C) / C)
        params += [np.asarray(v, np.float32) for v in fo.rbj_lowpass(f, 0.8, 48000.0)]
    pp = zg.compile(pexpr).plan(channels=C, mode=zg.MODE_EXACT)
    for j, v in enumerate(params):
        pp.set_param(j, v)
    yp = pp.process([zignal_b200.to_block(x[:, :T1])], n_samples=T1)[0].cpu().numpy()
    assert b"boxes per hand-over" in pp.info().kernel
    prm = np.stack(params, axis=1)
    assert np.array_equal(yp, fo.COracle(pexpr, C, params=prm).process([x[:, :T1]])[0])
    # FAST keeps K1b / the time segments; an explicit lanes_per_channel is respected
    assert b"split" not in k1b.info().kernel


def test_split_in_place(zg):
    C, T = 64, 2048
    x = fo.noise(C, T, seed=7)
    expr = fo.biquad_cascade(4)
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2, in_place=True)
    assert _is_split(plan)
    assert np.array_equal(y, fo.COracle(expr, C).process([x])[0])


def test_split_falls_back_when_the_shape_does_not_allow_it(zg):
    """T not a multiple of 32 samples: the lane-per-channel kernel runs, whatever section_warps says"""
    C, T = 64, 1000
    x = fo.noise(C, T, seed=8)
    expr = fo.biquad_cascade(4)
    y, plan = _run(zg, expr, x, zg.MODE_EXACT, 2)
    assert not _is_split(plan)
    assert np.array_equal(y, fo.COracle(expr, C).process([x])[0])


@pytest.mark.parametrize("sections", [3, 8])
def test_split_auto_other_section_counts_between_few_and_many_channels(zg, sections):
    """9600 channels = 300 channel groups: two groups per CTA on every SM; bit-identical to the lane-per-channel kernel"""
    torch = _torch()
    C, T = 9600, 4096
    expr = fo.biquad_cascade(sections)
    x = torch.rand((C, T), device="cuda") * 2 - 1
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT)
    y = plan.process([x], n_samples=T)[0]
    assert _is_split(plan), plan.info().kernel
    k1 = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, section_warps=1)
    assert torch.equal(y, k1.process([x], n_samples=T)[0]) and not _is_split(k1)
    idx = [0, 31, 4800, 9599]
    assert np.array_equal(y[idx].cpu().numpy(), fo.COracle(expr, len(idx)).process([x[idx].cpu().numpy()])[0])


def test_split_is_what_auto_picks_for_many_channels(zg):
    """the north-star shape class: 65 536 channels -> 2048 channel groups over 2 x SMs persistent groups"""
    torch = _torch()
    import zignal_b200
    C, T = 65536, 4096
    expr = fo.biquad_cascade(4)
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT)
    x = torch.rand((C, T), device="cuda") * 2 - 1
    y = plan.process([x], n_samples=T)[0]
    assert _is_split(plan), plan.info().kernel
    k1 = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, section_warps=1)
    z = k1.process([x], n_samples=T)[0]
    assert not _is_split(k1)
    assert torch.equal(y, z)
