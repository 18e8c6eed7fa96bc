"""Time-segmented launches for few, long channels (zg_plan_opts.time_parallel).

The reference evaluates a recurrence sample after sample (binary_feedback, flowz/flowz.hpp:1031-1074).  A linear tick
can be cut in time; both forms re-associate the arithmetic, so they exist in FAST mode only and are held to the FAST
bars of tests/test_gpu_parity.py, stated per graph:
  * well-conditioned graphs (one-pole low-pass, DF2 / DF1T forms with the reference's coefficients on short blocks):
    <= 1e-5 block-relative against the oracle;
  * the benchmark's RBJ cascade, whose own fp32 noise floor is ~1e-5 from float64
    (tests/test_oracle_golden.py::test_reference_rounding_noise_floor): no further from the float64 result than the
    reference is (median over channels) and <= 3e-5 from the reference -- the same bar the serial FAST kernel has;
  * and, because a cut in time must not add anything beyond rounding: within 4e-6 block-relative of the SERIAL FAST
    kernel (same FMA arithmetic, lane per channel, no segments) on well-conditioned graphs, and within two of its noise
    floors (2e-5) on the RBJ cascade -- a segment that starts from a warmed-up or fixed-up state carries a different
    realisation of the rounding noise the serial evaluation has accumulated by then (measured: 8.4e-6).
The host half (state matrix, settling time) is checked without a GPU against numpy.
"""
import numpy as np
import pytest

import flowz_oracle as fo
import reference_vectors as rv

TOL = 1e-5
FAST_TOL_BIQUAD = 3e-5
VS_SERIAL = 4e-6
VS_SERIAL_BIQUAD = 2e-5


def _rel_err(y, ref):
    den = np.abs(ref).max(axis=1)
    den = np.where(den == 0, 1.0, den)
    return (np.abs(y.astype(np.float64) - ref.astype(np.float64)).max(axis=1) / den).max()


# ---- host analysis (no GPU) -------------------------------------------------------------------------------------

def test_state_matrix_matches_the_probed_state_space(zg):
    g = zg.compile(fo.biquad_cascade(4))
    A = g.state_matrix()
    assert A.shape == (g.n_state, g.n_state)
    assert np.array_equal(A, g.state_space()[0])
    # poles of the four RBJ sections + the nilpotent input delay line
    ev = np.sort(np.abs(np.linalg.eigvals(A)))[::-1]
    r = [np.sqrt(-float(fo.rbj_lowpass(440.0 * 2 ** k)[4])) for k in range(4)]
    assert np.allclose(ev[:8], np.sort(np.repeat(r, 2))[::-1], atol=1e-6)


def test_settling_time_is_the_first_power_below_the_tolerance(zg):
    g = zg.compile(fo.biquad_cascade(4))
    A = g.state_matrix()
    K = g.settling_time(step=128, k_max=8192, tol=2.0 ** -30)
    assert K > 0 and K % 128 == 0
    norm = lambda M: np.abs(M).sum(axis=1).max()
    assert norm(np.linalg.matrix_power(A, K)) <= 2.0 ** -30 < norm(np.linalg.matrix_power(A, K - 128))
    # one-pole low-pass 0.9: 0.9^K <= 2^-30 from K = 198 on -> first multiple of 128
    assert zg.compile("~(_2 + 0.9f*_1[_1])").settling_time() == 256
    # an affine tick has the same A; an oscillator (poles on the unit circle) never forgets
    assert zg.compile("~(_2 + 0.9f*_1[_1] + 1.0f)").settling_time() == 256
    assert zg.compile(fo.osc_expr()).settling_time() == 0
    # per-parameter values
    gp = zg.compile("~(_2 + $0*_1[_1])")
    assert gp.settling_time([0.5]) == 128 and gp.settling_time([1.0]) == 0 and gp.settling_time([1.5]) == 0
    with pytest.raises(zg.ZgError):
        zg.compile("~(_2 + _1[_1]*_1[_1])").state_matrix()           # not linear
    with pytest.raises(zg.ZgError):
        gp.state_matrix([])                                         # one value per $k


def test_plan_options_are_validated_without_a_device(zg):
    g = zg.compile(fo.biquad_cascade(4))
    with pytest.raises(zg.ZgError) as e:
        g.plan(channels=64, mode=zg.MODE_FAST, time_parallel=7)
    assert e.value.status == zg.ZG_ERR_ARG


# ---- device ---------------------------------------------------------------------------------------------------------

def _plan_run(zg, expr, x, time_parallel, layout="planar", params=None, blocks=None, mode=None, input_kind=None, **kw):
    import torch
    g = zg.compile(expr)
    C, T = x[0].shape
    plan = g.plan(channels=C, mode=zg.MODE_FAST if mode is None else mode, time_parallel=time_parallel,
                  layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED, input_kind=input_kind, **kw)
    for i, p in enumerate(params or []):
        plan.set_param(i, p)
    outs = [[] for _ in range(g.n_out)]
    t0 = 0
    infos = []
    for n in (blocks or [T]):
        ins = []
        for k in range(g.n_in):
            if input_kind and input_kind[k] != zg.IN_BUFFER:
                ins.append(None)
                continue
            xb = x[k][:, t0:t0 + n]
            ins.append(zg.to_block(xb.T if layout == "interleaved" else xb))
        ys = plan.process(ins, n_samples=n)
        torch.cuda.synchronize()
        infos.append(plan.info())
        for o, y in zip(outs, ys):
            y = y.cpu().numpy()
            o.append(y.T if layout == "interleaved" else y)
        t0 += n
    return [np.concatenate(o, axis=1) for o in outs], plan, infos


def _oracle(expr, x, params=None):
    C = x[0].shape[0]
    prm = None
    if params is not None:
        prm = np.stack([np.broadcast_to(np.asarray(p, np.float32), (C,)) for p in params], axis=1)
    return fo.COracle(expr, C, params=prm).process(x)


def _check_biquad(y, ref, x, sections=4):
    truth = fo.biquad_cascade_f64(x, sections)
    den = np.abs(truth).max(axis=1)
    e_fast = np.abs(y - truth).max(axis=1) / den
    e_ref = np.abs(ref - truth).max(axis=1) / den
    assert np.median(e_fast) <= np.median(e_ref), (np.median(e_fast), np.median(e_ref))
    assert _rel_err(y, ref) <= FAST_TOL_BIQUAD


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["warmup", "two_pass"])
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_biquad_cascade_cut_in_time(zg, form, layout):
    C, T = 96, 12 * 1024 + 200                     # ragged: the last segment is shorter, the last box partial
    x = [fo.noise(C, T, seed=31)]
    expr = fo.biquad_cascade(4)
    tp = zg.TP_WARMUP if form == "warmup" else zg.TP_TWO_PASS
    ys, plan, infos = _plan_run(zg, expr, x, tp, layout)
    i = infos[0]
    assert i.time_segments >= 2 and i.lanes_per_channel == 1 and i.jit == 0
    assert (b"+segments:warm-up" if form == "warmup" else b"+segments:two-pass") in i.kernel
    assert (i.warmup_samples > 0) == (form == "warmup")
    assert i.launches == (1 if form == "warmup" else (2 if i.time_segments == 2 else 3))
    serial, _, si = _plan_run(zg, expr, x, zg.TP_OFF, layout, lanes_per_channel=1)
    assert si[0].time_segments == 1
    assert 0 < _rel_err(ys[0], serial[0]) <= VS_SERIAL_BIQUAD
    _check_biquad(ys[0], _oracle(expr, x)[0], x[0])


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["warmup", "two_pass"])
def test_cut_blocks_continue_the_stream(zg, form):
    """Consecutive zg_process calls continue like consecutive ticks: segment 0 starts from the plan's state, the last
    segment leaves it; the final state equals the serial kernel's to rounding."""
    C, T = 64, 3 * 6144
    x = [fo.noise(C, T, seed=32)]
    expr = fo.biquad_cascade(4)
    tp = zg.TP_WARMUP if form == "warmup" else zg.TP_TWO_PASS
    ys, plan, infos = _plan_run(zg, expr, x, tp, blocks=[6144, 6144 + 2048, 6144 - 2048])
    assert all(i.time_segments >= 2 for i in infos)
    serial, splan, _ = _plan_run(zg, expr, x, zg.TP_OFF, lanes_per_channel=1)
    assert _rel_err(ys[0], serial[0]) <= VS_SERIAL_BIQUAD
    st, ss = plan.get_state(), splan.get_state()
    assert np.abs(st - ss).max() <= VS_SERIAL_BIQUAD * np.abs(ss).max()
    _check_biquad(ys[0], _oracle(expr, x)[0], x[0])


WELL_CONDITIONED = [
    "~(_2 + 0.9f*_1[_1])",                                              # BASELINE configs[0] graph
    "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])",                       # two inputs
    "~( (_2 + 0.3f*_1[_1]) |= (0.5f*_1 + 0.25f*_1[_1]) )",
    "_1 |= (_1[_1] , 0.5f*_1[_2] + 0.25f , _1)",                        # three outputs, one of them affine
    "~(_2 + 0.5f*_1[_1] + 0.125f)",                                     # affine recurrence
]


@pytest.mark.gpu
@pytest.mark.parametrize("expr", WELL_CONDITIONED)
@pytest.mark.parametrize("form", ["warmup", "two_pass"])
def test_generated_kernels_cut_in_time(zg, expr, form):
    g = zg.compile(expr)
    C, T = 70, 8192 + 36
    x = [fo.noise(C, T, seed=40 + k) for k in range(g.n_in)]
    tp = zg.TP_WARMUP if form == "warmup" else zg.TP_TWO_PASS
    ys, plan, infos = _plan_run(zg, expr, x, tp)
    assert infos[0].time_segments >= 2 and infos[0].jit == 1
    ref = _oracle(expr, x)
    serial, _, _ = _plan_run(zg, expr, x, zg.TP_OFF)
    for y, r, sr in zip(ys, ref, serial):
        assert _rel_err(y, r) <= TOL
        assert _rel_err(y, sr) <= VS_SERIAL


@pytest.mark.gpu
def test_two_pass_handles_ticks_that_never_forget(zg):
    """Poles on the unit circle: the warm-up form refuses (nothing decays), the two-pass form is exact up to
    rounding.  y' = y + x (an integrator) and the reference's DF1 / DF2 benchmark graphs, whose literal
    coefficients put a pole at z = -1 (test/benchmark.cpp:18-23).  Its DF1T / DF2T graphs (poles at
    radius 0.894 with the same literals) forget within 256 samples and run in both forms."""
    C, T = 64, 4096
    x = [fo.noise(C, T, seed=50) * np.float32(0.01)]
    bg = rv.bench_graphs()
    for expr, forgets in [("~(_2 + _1[_1])", False), (bg[1], False), (bg[2], False), (bg[3], True), (bg[4], True)]:
        g = zg.compile(expr)
        assert (g.settling_time() > 0) == forgets
        serial, _, _ = _plan_run(zg, expr, x, zg.TP_OFF)
        ref = _oracle(expr, x)[0]
        if forgets:
            ys, plan, infos = _plan_run(zg, expr, x, zg.TP_WARMUP)
            assert infos[0].time_segments >= 2 and infos[0].warmup_samples == 256
            assert _rel_err(ys[0], serial[0]) <= VS_SERIAL and _rel_err(ys[0], ref) <= TOL
        else:
            with pytest.raises(zg.ZgError) as e:
                _plan_run(zg, expr, x, zg.TP_WARMUP)
            assert e.value.status == zg.ZG_ERR_UNSUPPORTED
        ys, plan, infos = _plan_run(zg, expr, x, zg.TP_TWO_PASS)
        assert infos[0].time_segments >= 2
        assert _rel_err(ys[0], serial[0]) <= 1e-5, expr
        assert _rel_err(ys[0], ref) <= 2e-5, expr


@pytest.mark.gpu
def test_per_channel_coefficients_cut_in_time(zg):
    """Every channel has its own A: the warm-up length is the worst channel's, the fix-up uses one A^L per channel."""
    C, T, S = 100, 16384, 4
    x = [fo.noise(C, T, seed=33)]
    params = []
    for k in range(S):
        per_ch = np.array([fo.rbj_lowpass(440.0 * 2 ** k * (1 + c / C)) for c in range(C)], np.float32)
        params += [per_ch[:, j].copy() for j in range(5)]
    expr = fo.biquad_cascade_params(S)
    ref = _oracle(expr, x, params)[0]
    serial, _, _ = _plan_run(zg, expr, x, zg.TP_OFF, params=params, lanes_per_channel=1)
    for tp in (zg.TP_WARMUP, zg.TP_TWO_PASS):
        ys, plan, infos = _plan_run(zg, expr, x, tp, params=params)
        assert infos[0].time_segments >= 2 and infos[0].uniform_params == 0
        assert _rel_err(ys[0], serial[0]) <= VS_SERIAL_BIQUAD
        assert _rel_err(ys[0], ref) <= FAST_TOL_BIQUAD
    # one channel made an integrator-like section (a1 = 2, a2 = -1: double pole at z = 1): no warm-up length exists
    params[3] = params[3].copy(); params[4] = params[4].copy()
    params[3][7], params[4][7] = 2.0, -1.0
    with pytest.raises(zg.ZgError) as e:
        _plan_run(zg, expr, x, zg.TP_WARMUP, params=params)
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


@pytest.mark.gpu
def test_exact_mode_and_nonlinear_graphs_stay_serial(zg):
    g = zg.compile(fo.biquad_cascade(4))
    with pytest.raises(zg.ZgError) as e:
        g.plan(channels=64, mode=zg.MODE_EXACT, time_parallel=zg.TP_WARMUP)
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED
    with pytest.raises(zg.ZgError):
        zg.compile("~(_2 + 0.5f*_1[_1]*_1[_1])").plan(channels=64, mode=zg.MODE_FAST, time_parallel=zg.TP_TWO_PASS)
    with pytest.raises(zg.ZgError):
        zg.compile(fo.fir_expr(fo.fir_taps(64))).plan(channels=64, mode=zg.MODE_FAST, time_parallel=zg.TP_TWO_PASS)
    # AUTO in EXACT mode: serial, bit-identical as ever
    C, T = 64, 16384
    x = [fo.noise(C, T, seed=34)]
    expr = fo.biquad_cascade(4)
    ys, plan, infos = _plan_run(zg, expr, x, zg.TP_AUTO, mode=zg.MODE_EXACT)
    assert infos[0].time_segments == 1
    assert np.array_equal(ys[0], _oracle(expr, x)[0])


@pytest.mark.gpu
def test_in_place_blocks(zg):
    """The warm-up of a segment re-reads samples its predecessor has overwritten by then: refused when asked for
    explicitly, AUTO falls back; the two-pass form reads only its own segment and runs in place."""
    import torch
    C, T = 64, 16384
    x = fo.noise(C, T, seed=35)
    expr = fo.biquad_cascade(4)
    g = zg.compile(expr)
    serial = g.plan(channels=C, mode=zg.MODE_FAST, time_parallel=zg.TP_OFF, lanes_per_channel=1).process([zg.to_block(x)])[0].cpu().numpy()
    buf = zg.to_block(x)
    with pytest.raises(zg.ZgError) as e:
        g.plan(channels=C, mode=zg.MODE_FAST, time_parallel=zg.TP_WARMUP).process([buf], outputs=[buf])
    assert e.value.status == zg.ZG_ERR_ARG
    for tp in (zg.TP_TWO_PASS, zg.TP_AUTO):
        buf = zg.to_block(x)
        plan = g.plan(channels=C, mode=zg.MODE_FAST, time_parallel=tp)
        plan.process([buf], outputs=[buf])
        torch.cuda.synchronize()
        assert _rel_err(buf.cpu().numpy(), serial) <= VS_SERIAL_BIQUAD
        assert (plan.info().time_segments >= 2) == (tp == zg.TP_TWO_PASS) or tp == zg.TP_AUTO


@pytest.mark.gpu
@pytest.mark.parametrize("C,T", [(96, 16384), (4000, 8192), (333, 12 * 1024 + 96)])
def test_section_split_kernel_cut_in_time(zg, C, T):
    """whole 32-sample boxes: the warm-up form runs on K1s with (channel group, segment) rows -- one group per SM for 96
    channels, three per CTA for 4000; ragged channel counts, a last segment shorter than the others; a second block
    continues the stream (the final state comes from the last segments, through a second state buffer)"""
    import torch
    x = [fo.noise(C, 2 * T, seed=C)]
    expr = fo.biquad_cascade(4)
    ys, plan, infos = _plan_run(zg, expr, x, zg.TP_WARMUP, blocks=[T, T])
    for i in infos:
        assert b"zg_biquad_df1_split" in i.kernel and b"+segments:warm-up" in i.kernel, i.kernel
        assert i.time_segments >= 2 and i.warmup_samples >= 640 and i.lanes_per_channel == 1
    serial, _, _ = _plan_run(zg, expr, x, zg.TP_OFF, lanes_per_channel=1)
    assert 0 < _rel_err(ys[0], serial[0]) <= VS_SERIAL_BIQUAD
    _check_biquad(ys[0], _oracle(expr, x)[0], x[0])


@pytest.mark.gpu
def test_full_size_config1_4096x65536_is_cut_automatically(zg):
    """BASELINE configs[1] in FAST mode: 128 channel groups x 8 segments instead of 128 warps (or K1b's 512)."""
    import torch
    C, T = 4096, 65536
    expr = fo.biquad_cascade(4)
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((C, T), generator=gen, device="cuda") * 2 - 1
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_FAST)
    y = plan.process([x])[0]
    torch.cuda.synchronize()
    i = plan.info()
    # (640 ticks forget the state; the segmented K1s rounds the warm-up to whole tiles)
    assert i.time_segments >= 4 and 640 <= i.warmup_samples <= 1024 and i.lanes_per_channel == 1 and i.launches == 1
    assert b"zg_biquad_df1_split" in i.kernel and b"+segments:warm-up" in i.kernel
    assert i.segment_samples >= 8 * 640
    serial = zg.compile(expr).plan(channels=C, mode=zg.MODE_FAST, time_parallel=zg.TP_OFF, lanes_per_channel=1).process([x])[0]
    den = serial.abs().amax(dim=1)
    assert float(((y - serial).abs().amax(dim=1) / den).max()) <= VS_SERIAL_BIQUAD
    idx = [0, 5, 2047, 4095]
    xs = x[idx].cpu().numpy()
    _check_biquad(y[idx].cpu().numpy(), _oracle(expr, [xs])[0], xs)
    # the section-parallel kernel is what EXACT mode (and an explicit lanes_per_channel) still gets
    assert zg.compile(expr).plan(channels=C, mode=zg.MODE_FAST, lanes_per_channel=4).info().lanes_per_channel == 4


@pytest.mark.gpu
def test_segments_with_bf16_storage_two_inputs_and_host_chunks(zg):
    """The segment logic lives in the streaming skeleton, so it composes with everything the skeleton does: bf16 sample
    storage, several input wires, and the row chunks of zg_process_host (each chunk is cut on its own)."""
    import torch
    expr = "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])"
    C, T = 96, 8192
    x = [fo.noise(C, T, seed=60), fo.noise(C, T, seed=61)]
    ref = _oracle(expr, x)[0]
    g = zg.compile(expr)
    for tp in (zg.TP_WARMUP, zg.TP_TWO_PASS):
        plan = g.plan(channels=C, mode=zg.MODE_FAST, time_parallel=tp)
        y = plan.process_host(x)[0]
        assert plan.info().time_segments >= 2
        assert _rel_err(y, ref) <= TOL
    xb = [torch.from_numpy(fo.bf16_round(a)).cuda().to(torch.bfloat16) for a in x]
    refb = _oracle(expr, [fo.bf16_round(a) for a in x])[0]
    plan = g.plan(channels=C, mode=zg.MODE_FAST, time_parallel=zg.TP_WARMUP, io_dtype=zg.BF16)
    yb = plan.process(xb)[0].float().cpu().numpy()
    assert plan.info().time_segments >= 2
    assert _rel_err(yb, refb) <= 2.0 ** -7                      # one bf16 ulp (8 bits of mantissa) block-relative


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_random_linear_graphs_two_pass_equals_serial(zg, seed):
    """Random LINEAR / AFFINE flowz graphs (series, parallel, fan-out, feedback, delays up to 3, one to three inputs):
    the two-pass form -- zero-state pass, boundary fix-up with A^L from the tick program, true-state pass -- must
    reproduce the serial FAST kernel on whatever state layout the lowering produced (several lines, shared lines,
    several outputs).  Graphs that blow up (poles outside the unit circle amplify rounding without bound) are skipped."""
    import random
    from test_fuzz_frontend import _gen
    rng = random.Random(8800 + seed)
    C, T = 40, 3000
    ran = 0
    for _ in range(400):
        if ran >= 6:
            break
        expr = _gen(rng, rng.randint(3, 5), rng.randint(1, 3), consts=["0.5f", "0.25f", "-0.75f", "0x1p-1f", "-0.5f", "0.125f"], ops="+-*")
        if "~" not in expr:
            continue
        try:
            g = zg.compile(expr)
        except zg.ZgError:
            continue
        if not g.all_f32 or g.n_in < 1 or g.n_out < 1 or g.n_state < 1 or g.n_state > 32 or g.linearity() == zg.NONLINEAR:
            continue
        x = [fo.noise(C, T, seed=500 * seed + 7 * ran + k) for k in range(g.n_in)]
        try:
            serial, _, _ = _plan_run(zg, expr, x, zg.TP_OFF)
            if not all(np.isfinite(s).all() and np.abs(s).max() < 1e4 for s in serial):
                continue
            ys, plan, infos = _plan_run(zg, expr, x, zg.TP_TWO_PASS)
        except zg.ZgError as e:
            if e.status == zg.ZG_ERR_UNSUPPORTED:
                continue
            raise AssertionError(f"{expr}: {e}")
        assert infos[0].time_segments >= 2, expr
        for y, s in zip(ys, serial):
            den = np.maximum(np.abs(s).max(axis=1), 1e-20)
            err = (np.abs(y.astype(np.float64) - s).max(axis=1) / den).max()
            assert err <= 2e-5, f"{expr}: {err}"
        ran += 1
    assert ran >= 4
