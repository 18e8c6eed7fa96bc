"""GPU parity: every result comes from zg_process()/zg_process_host() through the C ABI on cuda:0 and
is compared with the oracle (oracle/flowz_oracle.py, pinned by tests/test_oracle_*.py).

Bars: EXACT mode (the default, and what bench.py times) is bit-identical to the oracle,
i.e. inside the north star's 1e-5 with room to spare; delay indexing is bit-exact in both modes.
FAST mode (FMA contraction) rounds differently from the x86 non-FMA reference.  On well-conditioned
graphs it stays within TOL = 1e-5 block-relative (max_t|y - y_ref| <= TOL * max_t|y_ref| per channel).
The benchmark's biquad cascade is NOT well conditioned: its 440 Hz section amplifies rounding noise so
much that the reference itself sits up to 1.04e-5 (median 0.5e-5) from the float64 evaluation of the
same filter (tests/test_oracle_golden.py::test_reference_rounding_noise_floor), so no evaluator that
rounds differently can promise 1e-5 against it.  For that graph FAST is held to: no further from the
float64 result than the reference is (median over channels), and within FAST_TOL_BIQUAD = 3e-5 of the
reference.
"""
import numpy as np
import pytest

import flowz_oracle as fo
import reference_vectors as rv

pytestmark = pytest.mark.gpu

TOL = 1e-5
FAST_TOL_BIQUAD = 3e-5


def _check_fast_biquad(y, ref, x, sections):
    truth = fo.biquad_cascade_f64(x, sections)
    den = np.abs(truth).max(axis=1)
    e_fast = np.abs(y - truth).max(axis=1) / den
    e_ref = np.abs(ref - truth).max(axis=1) / den
    assert np.median(e_fast) <= np.median(e_ref), (np.median(e_fast), np.median(e_ref))
    err = _rel_err(y, ref)
    assert 0 < err <= FAST_TOL_BIQUAD, err           # 0 < : FMA contraction really is a different rounding


def _torch():
    import torch
    return torch


def _to_dev(x):
    import zignal_b200
    return zignal_b200.to_block(x)          # row pitch padded to a multiple of 4 floats when needed


def _run(zg, expr, x, mode, layout="planar", force_jit=False, params=None, blocks=None, input_kind=None, lanes=1):
    """x: [n_in][C, T] numpy.  Returns [n_out][C, T] numpy, plan."""
    torch = _torch()
    g = zg.compile(expr)
    C, T = x[0].shape if len(x) else (None, None)
    plan = g.plan(channels=C, mode=mode, layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED,
                  force_jit=force_jit, input_kind=input_kind, lanes_per_channel=lanes)
    for i, p in enumerate(params or []):
        plan.set_param(i, p)
    outs = [[] for _ in range(g.n_out)]
    t0 = 0
    for n in (blocks or [T]):
        ins = []
        for k in range(g.n_in):
            if input_kind and input_kind[k] != zg.IN_BUFFER:
                ins.append(None)
                continue
            xb = x[k][:, t0:t0 + n]
            ins.append(_to_dev(xb.T if layout == "interleaved" else xb))
        ys = plan.process(ins, n_samples=n)
        torch.cuda.synchronize()
        for o, y in zip(outs, ys):
            y = y.cpu().numpy()
            o.append(y.T if layout == "interleaved" else y)
        t0 += n
    return [np.concatenate(o, axis=1) for o in outs], plan


def _rel_err(y, ref):
    den = np.abs(ref).max(axis=1)
    den = np.where(den == 0, 1.0, den)
    return (np.abs(y.astype(np.float64) - ref.astype(np.float64)).max(axis=1) / den).max()


def _oracle(expr, x, params=None):
    C = x[0].shape[0]
    prm = None
    if params is not None:
        prm = np.stack([np.broadcast_to(np.asarray(p, np.float32), (C,)) for p in params], axis=1)
    return fo.COracle(expr, C, params=prm).process(x)


# ---- K1: prebuilt biquad cascade ---------------------------------------------------------------

@pytest.mark.parametrize("sections", [1, 2, 4, 8])
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_biquad_cascade_exact_is_bit_identical(zg, sections, layout):
    C, T = 96, 1000                      # ragged: C not a multiple of 32 lanes, T not of 32 samples... (T % 4 == 0)
    x = [fo.noise(C, T, seed=sections)]
    expr = fo.biquad_cascade(sections)
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, layout)
    info = plan.info()
    assert info.jit == 0 and b"zg_biquad_df1" in info.kernel and info.launches == 1
    ref = _oracle(expr, x)
    assert np.array_equal(ys[0], ref[0])


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_biquad_cascade_fast_within_tolerance(zg, layout):
    C, T = 200, 4096
    x = [fo.noise(C, T, seed=9)]
    expr = fo.biquad_cascade(4)
    ys, _ = _run(zg, expr, x, zg.MODE_FAST, layout)
    ref = _oracle(expr, x)
    _check_fast_biquad(ys[0], ref[0], x[0], 4)


def test_reference_benchmark_graph_df1_dirac(zg, ):
    """test/benchmark.cpp:157-167: DF1 graph, 201-sample dirac, reference coefficients."""
    x = np.zeros((32, 204), np.float32); x[:, 0] = 1.0
    ys, plan = _run(zg, rv.bench_graphs()[1], [x], zg.MODE_EXACT)
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "biquad_ref.npz"))
    assert plan.info().jit == 0
    for c in range(32):
        assert np.array_equal(ys[0][c, :201], g["custom1_dirac"])     # the reference's own hand-written loop


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_symmetric_and_asymmetric_numerators_take_different_ticks(zg, layout):
    """EXACT, shared coefficients: sections with b0 == b2 (bit for bit) reuse the product b0*x two ticks later
    instead of recomputing b2*x2 -- the same operands, so still bit-identical; any other coefficient set runs
    the nine-instruction tick.  Switching between them by zg_param_set keeps the stream continuous."""
    S, C, T = 4, 70, 900
    x = [fo.noise(C, T, seed=61)]
    expr = fo.biquad_cascade_params(S)
    sym = [v for k in range(S) for v in fo.rbj_lowpass(440.0 * 2 ** k)]
    asym = list(sym)
    asym[2] = np.float32(asym[2] * 1.25)                     # b2 != b0 in the first section
    for params, tag in ((sym, True), (asym, False)):
        ys, plan = _run(zg, expr, x, zg.MODE_EXACT, layout, params=[float(v) for v in params])
        assert (b"+b0=b2" in plan.info().kernel) == tag
        assert np.array_equal(ys[0], _oracle(expr, x, [float(v) for v in params])[0])
    # one stream, coefficients changed between blocks: symmetric -> asymmetric -> symmetric
    g = zg.compile(expr)
    plan = g.plan(channels=C, mode=zg.MODE_EXACT, layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED)
    outs, t0 = [], 0
    for params, n in ((sym, 300), (asym, 301), (sym, 299)):
        for i, v in enumerate(params):
            plan.set_param(i, float(v))
        xb = x[0][:, t0:t0 + n]
        yb = plan.process([_to_dev(xb.T if layout == "interleaved" else xb)])[0].cpu().numpy()
        outs.append(yb.T if layout == "interleaved" else yb)
        t0 += n
    # oracle: same schedule (state carried by hand through three oracles sharing a state vector is not exposed;
    # the C oracle takes per-channel parameters only per run, so compare block by block with explicit states)
    ref_plan = g.plan(channels=C, mode=zg.MODE_EXACT, force_jit=True)          # generated kernel: no product reuse
    t0 = 0
    for k, (params, n) in enumerate(((sym, 300), (asym, 301), (sym, 299))):
        for i, v in enumerate(params):
            ref_plan.set_param(i, float(v))
        yb = ref_plan.process([_to_dev(x[0][:, t0:t0 + n])])[0].cpu().numpy()
        assert np.array_equal(outs[k], yb)
        t0 += n


def test_per_channel_coefficients(zg):
    C, T, S = 128, 512, 2
    x = [fo.noise(C, T, seed=4)]
    params = []
    for k in range(S):
        per_ch = np.array([fo.rbj_lowpass(440.0 * 2 ** k * (1 + c / C)) for c in range(C)], np.float32)   # [C, 5]
        params += [per_ch[:, j].copy() for j in range(5)]
    expr = fo.biquad_cascade_params(S)
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, params=params)
    assert plan.info().uniform_params == 0
    assert np.array_equal(ys[0], _oracle(expr, x, params)[0])
    # scalar parameters take the constant-bank variant and still match
    scal = [float(p[0]) for p in params]
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, params=scal)
    assert plan.info().uniform_params == 1
    assert np.array_equal(ys[0], _oracle(expr, x, scal)[0])


def test_streaming_blocks_equal_one_block(zg):
    C, T = 64, 2048
    x = [fo.noise(C, T, seed=6)]
    expr = fo.biquad_cascade(4)
    whole, _ = _run(zg, expr, x, zg.MODE_EXACT)
    parts, plan = _run(zg, expr, x, zg.MODE_EXACT, blocks=[4, 31, 29, 1, 999, 984])
    assert plan.info().launches == 6
    assert np.array_equal(whole[0], parts[0])


def test_state_get_set_reset(zg):
    torch = _torch()
    C, T = 40, 256
    x = fo.noise(C, 2 * T, seed=8)
    g = zg.compile(fo.biquad_cascade(2))
    p1 = g.plan(channels=C, mode=zg.MODE_EXACT)
    y_full = torch.cat([p1.process([_to_dev(x[:, :T])])[0], p1.process([_to_dev(x[:, T:])])[0]], dim=1).cpu().numpy()
    p2 = g.plan(channels=C, mode=zg.MODE_EXACT)
    p2.process([_to_dev(x[:, :T])])
    st = p2.get_state()
    assert st.shape == (g.n_state, C) and np.abs(st).max() > 0
    p3 = g.plan(channels=C, mode=zg.MODE_EXACT)          # "copying the callable copies its state"
    p3.set_state(st)
    y3 = p3.process([_to_dev(x[:, T:])])[0].cpu().numpy()
    assert np.array_equal(y3, y_full[:, T:])
    p3.reset()
    assert np.abs(p3.get_state()).max() == 0
    assert np.array_equal(p3.process([_to_dev(x[:, :T])])[0].cpu().numpy(), y_full[:, :T])


# ---- K1b: section-parallel biquad cascade (S lanes per channel) ---------------------------------

@pytest.mark.parametrize("sections", [2, 4])
@pytest.mark.parametrize("C,T", [(1, 4), (8, 31), (9, 64), (70, 257), (96, 1000), (33, 4100), (520, 8192),
                                 (40, 1536), (17, 1537), (20000, 2100)])
def test_biquad_lanes_exact_is_bit_identical(zg, sections, C, T):
    x = [fo.noise(C, T, seed=100 + sections)]
    expr = fo.biquad_cascade(sections)
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, lanes=sections)
    info = plan.info()
    assert info.lanes_per_channel == sections and b"zg_biquad_df1_lanes" in info.kernel and info.launches == 1
    assert np.array_equal(ys[0], _oracle(expr, x)[0])


def test_biquad_lanes_auto_selected_for_few_channels(zg):
    g = zg.compile(fo.biquad_cascade(4))
    assert g.plan(channels=4096).info().lanes_per_channel == 4          # BASELINE configs[1]
    assert g.plan(channels=65536).info().lanes_per_channel == 1         # north-star shape
    assert g.plan(channels=4096, layout=zg.INTERLEAVED).info().lanes_per_channel == 1
    assert zg.compile(fo.biquad_cascade(3)).plan(channels=64).info().lanes_per_channel == 1
    with pytest.raises(zg.ZgError) as e:
        zg.compile(fo.biquad_cascade(3)).plan(channels=64, lanes_per_channel=4)
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


def test_biquad_lanes_streaming_and_state_interchange(zg):
    """ragged blocks continue bit for bit; the state rows written by one kernel are read by the other"""
    C, T = 40, 3000
    x = [fo.noise(C, T, seed=77)]
    expr = fo.biquad_cascade(4)
    ref = _oracle(expr, x)[0]
    parts, plan = _run(zg, expr, x, zg.MODE_EXACT, lanes=4, blocks=[5, 1, 250, 7, 1024, 1713])
    assert plan.info().launches == 6
    assert np.array_equal(parts[0], ref)
    g = zg.compile(expr)
    a = g.plan(channels=C, lanes_per_channel=4)
    b = g.plan(channels=C, lanes_per_channel=1)
    y1 = a.process([_to_dev(x[0][:, :1001])])[0].cpu().numpy()
    b.set_state(a.get_state())
    y2 = b.process([_to_dev(x[0][:, 1001:2000])])[0].cpu().numpy()
    a.set_state(b.get_state())
    y3 = a.process([_to_dev(x[0][:, 2000:])])[0].cpu().numpy()
    assert np.array_equal(np.concatenate([y1, y2, y3], axis=1), ref)


def test_biquad_lanes_per_channel_coefficients_and_fast_mode(zg):
    C, T, S = 100, 2048, 4
    x = [fo.noise(C, T, seed=5)]
    params = []
    for k in range(S):
        per_ch = np.array([fo.rbj_lowpass(440.0 * 2 ** k * (1 + c / C)) for c in range(C)], np.float32)
        params += [per_ch[:, j].copy() for j in range(5)]
    expr = fo.biquad_cascade_params(S)
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, params=params, lanes=4)
    assert plan.info().uniform_params == 0 and plan.info().lanes_per_channel == 4
    assert np.array_equal(ys[0], _oracle(expr, x, params)[0])
    # FMA mode: same association as the lane-per-channel kernel -> the two agree bit for bit
    e4 = fo.biquad_cascade(4)
    f4, _ = _run(zg, e4, x, zg.MODE_FAST, lanes=4)
    f1, _ = _run(zg, e4, x, zg.MODE_FAST, lanes=1)
    assert np.array_equal(f4[0], f1[0])
    _check_fast_biquad(f4[0], _oracle(e4, x)[0], x[0], 4)


# ---- K2: generated tick (NVRTC) -----------------------------------------------------------------

GENERIC = [
    "~(_2 + 0.9f*_1[_1])",                                             # config 1 graph
    "~(0x1.fcp0f*_1[_1] - _1[_2] + _2) |= ~(_2 + 0.9f*_1[_1])",        # config 3: osc >> one-pole
    "_1 |= (_1[_1] , _1[_3]) |= _1 - _2",
    "(_1 , _1[_2]) |= (_1 | 0.5f*_1) |= _1*_2",
    "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])",                      # two inputs
    "~( (_2 + 0.3f*_1[_1]) |= (0.5f*_1 + 0.25f*_1[_1]) )",
    "_1 |= (_1[_1] , _1[_2] , _1)",                                    # three outputs
    "_1 / (_1[_1]*_1[_1] + 1.5f) - -_1[_2]",
    rv.bench_graphs()[2], rv.bench_graphs()[3], rv.bench_graphs()[4],  # DF2, DF1T, DF2T of the reference
]


@pytest.mark.parametrize("expr", GENERIC)
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_generated_kernel_exact_is_bit_identical(zg, expr, layout):
    g = zg.compile(expr)
    C, T = 70, 612
    x = [fo.noise(C, T, seed=20 + k) for k in range(g.n_in)]
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, layout)
    assert plan.info().jit == 1
    ref = _oracle(expr, x)
    for y, r in zip(ys, ref):
        assert np.array_equal(y, r, equal_nan=True)


# GENERIC[1] (the oscillator) is left out on purpose: its poles sit ON the unit circle, rounding
# differences grow linearly with time, and FMA-vs-non-FMA exceeds 1e-5 within ~1000 samples (measured
# 2e-5 at 2048).  Marginally stable graphs have to run in EXACT mode.
@pytest.mark.parametrize("expr", GENERIC[:1] + GENERIC[2:6] + GENERIC[8:])
def test_generated_kernel_fast_within_tolerance(zg, expr):
    g = zg.compile(expr)
    C, T = 64, 2048
    x = [fo.noise(C, T, seed=30 + k) for k in range(g.n_in)]
    ys, _ = _run(zg, expr, x, zg.MODE_FAST)
    ref = _oracle(expr, x)
    for y, r in zip(ys, ref):
        assert _rel_err(y, r) <= TOL


def test_generated_equals_prebuilt_for_biquads(zg):
    x = [fo.noise(64, 1024, seed=2)]
    expr = fo.biquad_cascade(4)
    for mode in (zg.MODE_EXACT, zg.MODE_FAST):
        a, pa = _run(zg, expr, x, mode)
        b, pb = _run(zg, expr, x, mode, force_jit=True)
        assert pa.info().jit == 0 and pb.info().jit == 1
        if mode == zg.MODE_EXACT:
            assert np.array_equal(a[0], b[0])
        else:
            assert _rel_err(a[0], b[0]) <= FAST_TOL_BIQUAD


def test_delay_indexing_is_bit_exact_on_integer_ramp(zg):
    """pure delays: an integer-valued ramp must come back shifted, bit for bit, in every mode."""
    C, T = 33, 300
    ramp = (np.arange(T, dtype=np.float32)[None, :] + 1000 * np.arange(C, dtype=np.float32)[:, None])
    expr = "_1 |= (_1[_1] , _1[_5] , _1[_17])"
    for mode in (zg.MODE_EXACT, zg.MODE_FAST):
        for layout in ("planar", "interleaved"):
            ys, _ = _run(zg, expr, [ramp], mode, layout, blocks=[100, 8, 192])
            for y, n in zip(ys, (1, 5, 17)):
                want = np.zeros_like(ramp); want[:, n:] = ramp[:, :-n]
                assert np.array_equal(y, want)


def test_dirac_excited_oscillator_synthesised_input(zg):
    """config 3: the excitation is synthesised in the kernel (no input buffer is read)."""
    C, T = 64, 1024
    k = np.float32(2 * np.cos(2 * np.pi * 440.0 / 44100.0))
    expr = f"~({fo.lit(k)}*_1[_1] - _1[_2] + _2) |= ~(_2 + 0.9f*_1[_1])"
    x = np.zeros((C, T), np.float32); x[:, 0] = 1.0
    ys, plan = _run(zg, expr, [x], zg.MODE_EXACT, input_kind=[zg.IN_DIRAC], blocks=[512, 512])
    ref = _oracle(expr, [x])
    assert np.array_equal(ys[0], ref[0])
    # zero input: stays silent
    ys, _ = _run(zg, expr, [x], zg.MODE_EXACT, input_kind=[zg.IN_ZERO])
    assert np.abs(ys[0]).max() == 0


def test_per_channel_parameter_generated_kernel(zg):
    C, T = 96, 400
    x = [fo.noise(C, T, seed=12)]
    a = np.linspace(0.1, 0.95, C).astype(np.float32)
    expr = "~(_2 + $0*_1[_1])"
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, params=[a])
    assert np.array_equal(ys[0], _oracle(expr, x, [a])[0])


# ---- host-buffer entry point, edge cases, errors ---------------------------------------------------

def test_process_host_round_trip(zg):
    C, T = 50, 777                       # T not a multiple of 4: host rows are re-pitched on the device
    x = fo.noise(C, T, seed=3)
    expr = fo.biquad_cascade(4)
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT)
    y = plan.process_host([x])[0]
    assert np.array_equal(y, _oracle(expr, [x])[0])


@pytest.mark.parametrize("C,T", [(1, 4), (1, 1024), (31, 36), (32, 32), (33, 4100), (4096, 64)])
def test_shapes(zg, C, T):
    x = [fo.noise(C, T, seed=C + T)]
    expr = fo.biquad_cascade(2)
    ys, _ = _run(zg, expr, x, zg.MODE_EXACT)
    assert np.array_equal(ys[0], _oracle(expr, x)[0])


def test_empty_block_is_a_no_op(zg):
    plan = zg.compile(fo.biquad_cascade(1)).plan(channels=8)
    plan.process_ptrs([16], [16], 0, 4, 4)
    assert plan.info().launches == 0


def test_argument_errors(zg):
    torch = _torch()
    g = zg.compile(fo.biquad_cascade(1))
    plan = g.plan(channels=8)
    x = torch.zeros(8, 64, device="cuda")
    with pytest.raises(zg.ZgError) as e:
        plan.process_ptrs([x.data_ptr() + 4], [x.data_ptr()], 64, 64, 64)     # misaligned
    assert e.value.status == zg.ZG_ERR_ARG
    with pytest.raises(zg.ZgError):
        plan.process_ptrs([x.data_ptr()], [x.data_ptr()], 64, 62, 64)         # ld too small
    with pytest.raises(zg.ZgError) as e:
        zg.compile("_1 + 1").plan(channels=8)                                  # int terminal: host only
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


# ---- K3: dense FIR (BASELINE configs[3]) -----------------------------------------------------------

@pytest.mark.parametrize("n_taps,C,T", [(256, 64, 1024), (256, 33, 700), (2, 32, 64), (17, 40, 100), (33, 5, 31),
                                        (100, 64, 4096), (255, 32, 513), (257, 32, 640), (512, 96, 2048)])
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_fir_exact_is_bit_identical(zg, n_taps, C, T, layout):
    h = fo.fir_taps(n_taps)
    expr = fo.fir_expr(h)
    x = fo.noise(C, T, seed=n_taps + C)
    ys, plan = _run(zg, expr, [x], zg.MODE_EXACT, layout)
    assert plan.info().kernel.decode() == f"zg_fir<{n_taps} taps,exact,{layout}>" and plan.info().jit == 0
    assert np.array_equal(ys[0], fo.fir_direct(x, h))
    if C * T <= 64 * 1024:                                   # and the tick-by-tick oracle itself
        assert np.array_equal(ys[0], _oracle(expr, [x])[0])


def test_fir_fast_within_tolerance(zg):
    h = fo.fir_taps(256)
    x = fo.noise(64, 2048, seed=5)
    ys, _ = _run(zg, fo.fir_expr(h), [x], zg.MODE_FAST)
    ref = fo.fir_direct(x, h)
    err = _rel_err(ys[0], ref)
    assert 0 < err <= TOL, err


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_fir_streaming_blocks_and_state(zg, layout):
    """Ragged block lengths (shorter than the delay line, not multiples of the 32-sample box) continue
    exactly like consecutive ticks; the delay line is visible through zg_state_get in the oracle's
    slot order (oldest first, rotate_push_back flowz.hpp:130-148)."""
    h = fo.fir_taps(256)
    expr = fo.fir_expr(h)
    C, T = 40, 1500
    x = fo.noise(C, T, seed=9)
    ys, plan = _run(zg, expr, [x], zg.MODE_EXACT, layout, blocks=[7, 100, 33, 255, 256, 1, 848])
    assert np.array_equal(ys[0], fo.fir_direct(x, h))
    st = plan.get_state()                                     # [255][C]
    assert np.array_equal(st, x[:, -255:].T)
    # set_state: a plan started from that delay line continues the stream
    x2 = fo.noise(C, 300, seed=10)
    inter = layout == "interleaved"
    plan2 = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, layout=zg.INTERLEAVED if inter else zg.PLANAR)
    plan2.set_state(st)
    run2 = lambda: (lambda y: y.T if inter else y)(plan2.process([_to_dev(x2.T if inter else x2)])[0].cpu().numpy())
    assert np.array_equal(run2(), fo.fir_direct(x2, h, history=x[:, -255:]))
    plan2.reset()
    assert np.array_equal(run2(), fo.fir_direct(x2, h))


def test_fir_time_segments_for_few_channels(zg):
    """Few channel groups: the block is cut into time segments (a FIR has no recurrence); the segment
    that ends the block writes the delay line."""
    h = fo.fir_taps(256)
    C, T = 64, 16384
    x = fo.noise(C, T, seed=11)
    ys, plan = _run(zg, fo.fir_expr(h), [x], zg.MODE_EXACT, blocks=[T // 2, T // 2])
    assert plan.info().launches == 2
    assert np.array_equal(ys[0], fo.fir_direct(x, h))
    assert np.array_equal(plan.get_state(), x[:, -255:].T)


def test_fir_taps_as_parameters(zg):
    n = 48
    h = fo.fir_taps(n)
    x = fo.noise(64, 512, seed=12)
    ys, _ = _run(zg, fo.fir_expr_params(n), [x], zg.MODE_EXACT, params=[float(v) for v in h])
    assert np.array_equal(ys[0], fo.fir_direct(x, h))
    g = zg.compile(fo.fir_expr_params(n))
    plan = g.plan(channels=64)
    plan.set_param(3, np.linspace(0, 1, 64, dtype=np.float32))      # per-channel tap: not this kernel
    with pytest.raises(zg.ZgError) as e:
        plan.process([_to_dev(x)])
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


def test_fir_delay_indexing_integer_ramp(zg):
    """Unit taps on an integer ramp: every output is an exactly representable integer sum."""
    n = 40
    x = np.tile(np.arange(1, 201, dtype=np.float32), (32, 1))
    ys, _ = _run(zg, fo.fir_expr(np.ones(n, np.float32)), [x], zg.MODE_FAST)
    c = np.concatenate([np.zeros(1), np.cumsum(x[0].astype(np.float64))])
    t = np.arange(200)
    want = c[t + 1] - c[np.maximum(t + 1 - n, 0)]
    assert np.array_equal(ys[0], np.tile(want.astype(np.float32), (32, 1)))


def test_fir_process_host_and_long_delay_errors(zg):
    h = fo.fir_taps(64)
    x = fo.noise(2048, 4096, seed=13)
    plan = zg.compile(fo.fir_expr(h)).plan(channels=2048)
    y = plan.process_host([x])[0]
    assert np.array_equal(y, fo.fir_direct(x, h))
    # a long delay line that is not a dense FIR runs the generated kernel (the line is a ring in HBM)
    ys, plan2 = _run(zg, "_1[_100] + 0.5f*_1", [x[:64, :600]], zg.MODE_EXACT)
    assert plan2.info().jit == 1 and np.array_equal(ys[0], _oracle("_1[_100] + 0.5f*_1", [x[:64, :600]])[0])
    # interleaved frames through the host path: chunks are time ranges, the delay line ping-pongs per chunk
    xi = np.ascontiguousarray(x.T)
    yi = zg.compile(fo.fir_expr(h)).plan(channels=2048, layout=zg.INTERLEAVED).process_host([xi])[0]
    assert np.array_equal(yi.T, fo.fir_direct(x, h))


def test_full_size_config4_fir256_32768_channels(zg):
    """BASELINE configs[3] at full size: 32 768 channels x 256 taps (T = 2048 here keeps the host-side
    check quick; the kernel path is the same for any T).  Properties: sampled channels bit-identical
    to the oracle, identical inputs give identical channels, two half blocks equal one block,
    linearity in FAST mode."""
    torch = _torch()
    C, T = 32768, 2048
    h = fo.fir_taps(256)
    g = zg.compile(fo.fir_expr(h))
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((C, T), generator=gen, device="cuda") * 2 - 1
    x[1::2] = x[0::2]
    plan = g.plan(channels=C, mode=zg.MODE_EXACT)
    y = plan.process([x])[0]
    torch.cuda.synchronize()
    assert torch.equal(y[0::2], y[1::2])
    idx = [0, 1, 31, 32, 4097, 32766, 32767]
    assert np.array_equal(y[idx].cpu().numpy(), fo.fir_direct(x[idx].cpu().numpy(), h))
    plan2 = g.plan(channels=C, mode=zg.MODE_EXACT)
    ya = plan2.process([x[:, :T // 2]])[0]
    yb = plan2.process([x[:, T // 2:]])[0]
    assert torch.equal(torch.cat([ya, yb], dim=1), y)
    yf = g.plan(channels=C, mode=zg.MODE_FAST).process([x])[0]
    assert _rel_err(yf[idx].cpu().numpy(), y[idx].cpu().numpy()) <= TOL


# ---- bf16 sample storage (BASELINE configs[4]) ------------------------------------------------------
# Bar: compare after rounding the oracle's fp32 output to bf16, <= 1 bf16 ulp.  EXACT mode
# does better: inputs are bf16-representable, arithmetic and state are fp32 exactly as the oracle's, the
# store rounds to nearest even -> the stored 16 bits are identical.

def _run_bf16(zg, expr, x, mode, layout="planar", blocks=None, input_kind=None, n_samples=None, channels=None):
    """x: [n_in][C, T] fp32 (rounded to bf16 on the way in).  Returns [n_out] uint16 arrays [C, T]."""
    torch = _torch()
    import zignal_b200
    g = zg.compile(expr)
    C, T = (x[0].shape if x and x[0] is not None else (channels, n_samples))
    plan = g.plan(channels=C, mode=mode, layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED,
                  io_dtype=zg.BF16, input_kind=input_kind)
    assert b"bf16" in plan.info().kernel
    outs = [[] for _ in range(g.n_out)]
    t0 = 0
    for n in (blocks or [T]):
        ins = []
        for k in range(g.n_in):
            if input_kind and input_kind[k] != zg.IN_BUFFER:
                ins.append(None)
                continue
            xb = x[k][:, t0:t0 + n]
            ins.append(zignal_b200.to_block(xb.T if layout == "interleaved" else xb, dtype=torch.bfloat16))
        ys = plan.process(ins, n_samples=n)
        torch.cuda.synchronize()
        for o, y in zip(outs, ys):
            y = y.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
            o.append(y.T if layout == "interleaved" else y)
        t0 += n
    return [np.concatenate(o, axis=1) for o in outs], plan


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
@pytest.mark.parametrize("C,T", [(64, 1024), (33, 100), (1, 7), (40, 4100)])
def test_bf16_storage_exact_mode_stores_identical_bits(zg, layout, C, T):
    expr = fo.biquad_cascade(2)
    x = fo.noise(C, T, seed=C * 7 + T)
    ys, _ = _run_bf16(zg, expr, [x], zg.MODE_EXACT, layout)
    ref = _oracle(expr, [fo.bf16_round(x)])[0]
    assert np.array_equal(ys[0], fo.bf16_bits(ref))


@pytest.mark.parametrize("expr", ["(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])",          # two inputs
                                  "_1 |= (_1[_1] , 0.5f*_1[_2] , _1)"])                    # three outputs
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_bf16_storage_streaming_and_multi_wire(zg, expr, layout):
    g = zg.compile(expr)
    x = [fo.noise(48, 900, seed=21 + k) for k in range(g.n_in)]
    ys, _ = _run_bf16(zg, expr, x, zg.MODE_EXACT, layout, blocks=[64, 1, 100, 735])
    ref = _oracle(expr, [fo.bf16_round(v) for v in x])
    assert len(ys) == g.n_out
    for y, r in zip(ys, ref):
        assert np.array_equal(y, fo.bf16_bits(r))


def test_bf16_polyphonic_voice_chain(zg):
    """BASELINE configs[4]'s graph: osc >> biquad >> (biquad in a feedback loop), dirac-excited (input
    synthesised in the kernel: no HBM read), bf16 output.  FAST mode within 1 bf16 ulp."""
    expr = fo.poly_voice_expr()
    C, T = 256, 2048
    d = np.zeros((C, T), np.float32); d[:, 0] = 1
    ref = _oracle(expr, [d])[0]
    ys, plan = _run_bf16(zg, expr, [None], zg.MODE_EXACT, input_kind=[zg.IN_DIRAC], n_samples=T, channels=C,
                         blocks=[1000, 1048])
    assert np.array_equal(ys[0], fo.bf16_bits(ref))
    yf, _ = _run_bf16(zg, expr, [None], zg.MODE_FAST, input_kind=[zg.IN_DIRAC], n_samples=T, channels=C)
    got = (yf[0].astype(np.uint32) << 16).view(np.float32)
    # FAST rounds differently, and the oscillator is marginally stable: its error is relative
    # to the block's amplitude, not to each sample (near a zero crossing one ulp of the sample is tiny).
    # Bar: within one bf16 ulp of the block's largest value, and most stored samples identical.
    ulp_max = 2.0 ** (np.floor(np.log2(np.abs(ref).max())) - 7)
    assert np.abs(got - fo.bf16_round(ref)).max() <= ulp_max
    assert np.mean(yf[0] == fo.bf16_bits(ref)) > 0.9


def test_bf16_process_host(zg):
    torch = _torch()
    expr = fo.biquad_cascade(4)
    C, T = 512, 1000
    x = fo.noise(C, T, seed=31)
    xh = torch.from_numpy(x).to(torch.bfloat16)
    y = zg.compile(expr).plan(channels=C, io_dtype=zg.BF16).process_host([xh])[0]
    ref = _oracle(expr, [fo.bf16_round(x)])[0]
    assert np.array_equal(y.view(torch.int16).numpy().view(np.uint16), fo.bf16_bits(ref))


def test_bf16_unsupported_kernels_say_so(zg):
    with pytest.raises(zg.ZgError) as e:
        zg.compile(fo.biquad_cascade(4)).plan(channels=64, io_dtype=zg.BF16, lanes_per_channel=4)
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED
    with pytest.raises(zg.ZgError) as e:
        zg.compile(fo.fir_expr(fo.fir_taps(256))).plan(channels=64, io_dtype=zg.BF16)
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


# ---- size-independent properties at BASELINE sizes ---------------------------------------------------

def test_full_size_properties_65536_channels(zg):
    """NS shape (65 536 ch; T shortened to keep the test quick on the host side is NOT done: the
    full 8192 samples run on the GPU, only the oracle is sampled).  Properties: (1) a sample of
    channels equals the oracle bit for bit, (2) identical inputs give identical channels,
    (3) two half blocks equal one block."""
    torch = _torch()
    C, T = 65536, 8192
    expr = fo.biquad_cascade(4)
    g = zg.compile(expr)
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((C, T), generator=gen, device="cuda") * 2 - 1
    x[1::2] = x[0::2]                                            # channel pairs share their input
    plan = g.plan(channels=C, mode=zg.MODE_EXACT)
    y = plan.process([x])[0]
    torch.cuda.synchronize()
    assert torch.equal(y[0::2], y[1::2])
    idx = [0, 1, 31, 32, 4097, 65534, 65535]
    ref = _oracle(expr, [x[idx].cpu().numpy()])[0]
    assert np.array_equal(y[idx].cpu().numpy(), ref)
    plan2 = g.plan(channels=C, mode=zg.MODE_EXACT)
    ya = plan2.process([x[:, :T // 2]])[0]
    yb = plan2.process([x[:, T // 2:]])[0]
    assert torch.equal(torch.cat([ya, yb], dim=1), y)
    # fast mode on the same data, sampled channels
    yf = g.plan(channels=C, mode=zg.MODE_FAST).process([x])[0]
    _check_fast_biquad(yf[idx].cpu().numpy(), ref, x[idx].cpu().numpy(), 4)


def test_full_size_config2_4096x65536(zg):
    torch = _torch()
    C, T = 4096, 65536
    expr = fo.biquad_cascade(4)
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((C, T), generator=gen, device="cuda") * 2 - 1
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT)
    y = plan.process([x])[0]
    # too few channels for a lane each, EXACT, long block: a warp per section, one group per SM (K1s, few-channel form)
    assert plan.info().lanes_per_channel == 1 and b"zg_biquad_df1_split<4,exact,planar,4 warps per group,4 boxes per hand-over>" in plan.info().kernel
    # ... bit-identical to the section-parallel lanes kernel (K1b) it replaced on this shape
    k1b = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=4)
    z = k1b.process([x], n_samples=T)[0]
    assert k1b.info().lanes_per_channel == 4 and torch.equal(y, z)
    idx = [0, 5, 4095]
    assert np.array_equal(y[idx].cpu().numpy(), _oracle(expr, [x[idx].cpu().numpy()])[0])
    y1 = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=1).process([x])[0]
    assert torch.equal(y, y1)                            # both kernels, every channel, bit for bit


def test_full_size_config5_one_million_voices_bf16(zg):
    """BASELINE configs[4] at its full voice count on one GPU (1 048 576 voices x 512 samples, bf16 out, dirac
    input synthesised in the kernel).  Every voice runs the same graph from the same excitation, so all rows
    must be identical, and equal to the oracle's row bit for bit; a second block continues the stream."""
    torch = _torch()
    C, T = 1 << 20, 512
    expr = fo.poly_voice_expr()
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, io_dtype=zg.BF16, input_kind=[zg.IN_DIRAC])
    y1 = plan.process([None], n_samples=T)[0]
    y2 = plan.process([None], n_samples=T)[0]
    torch.cuda.synchronize()
    y = torch.cat([y1, y2], dim=1).view(torch.int16)
    assert bool((y == y[0:1]).all())
    d = np.zeros((1, 2 * T), np.float32); d[0, 0] = 1
    ref = fo.bf16_bits(_oracle(expr, [d])[0])
    for c in (0, 12345, C - 1):
        assert np.array_equal(y[c].cpu().numpy().view(np.uint16), ref[0])


def test_sample_rate_parameter_is_just_another_input_wire(zg):
    """block-rate parameters are `$k` (zg_param_set); a parameter that changes every sample
    is expressed the way the reference would -- as one more input of the graph (here the feedback gain of a
    one-pole low-pass: y = x + a(t) * y1)."""
    expr = "~(_2 + _3*_1[_1])"
    g = zg.compile(expr)
    assert (g.n_in, g.n_out) == (2, 1)
    C, T = 96, 700
    x = fo.noise(C, T, seed=41)
    a = (0.5 + 0.45 * fo.noise(C, T, seed=42)).astype(np.float32)          # 0.05 .. 0.95, per channel and sample
    for mode in (zg.MODE_EXACT, zg.MODE_FAST):
        ys, _ = _run(zg, expr, [x, a], mode, blocks=[100, 600])
        ref = _oracle(expr, [x, a])[0]
        if mode == zg.MODE_EXACT:
            assert np.array_equal(ys[0], ref)
        else:
            assert _rel_err(ys[0], ref) <= TOL


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_mixed_buffer_and_synthesised_inputs(zg, layout):
    """Two inputs, one streamed from HBM and one synthesised in the kernel (dirac / zeros): only the buffer
    wire gets a TMA pipeline; blocks continue across the dirac's position."""
    expr = "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])"
    g = zg.compile(expr)
    assert g.n_in == 2
    C, T = 45, 500
    x0 = fo.noise(C, T, seed=71)
    d = np.zeros((C, T), np.float32); d[:, 0] = 1
    for kind, second in ((zg.IN_DIRAC, d), (zg.IN_ZERO, np.zeros_like(d))):
        ys, _ = _run(zg, expr, [x0, second], zg.MODE_EXACT, layout, input_kind=[zg.IN_BUFFER, kind], blocks=[1, 63, 436])
        assert np.array_equal(ys[0], _oracle(expr, [x0, second])[0])


def test_full_size_config3_65536_voice_oscillators(zg):
    """BASELINE configs[2] at full size: 65 536 voices of osc >> one-pole low-pass, dirac-excited, 16 384 samples.
    All voices are identical, and equal to the oracle bit for bit (EXACT: a marginally stable oscillator must not
    be re-rounded); two half blocks continue the stream."""
    torch = _torch()
    C, T = 65536, 16384
    expr = fo.osc_lp_expr()
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, input_kind=[zg.IN_DIRAC])
    ya = plan.process([None], n_samples=T // 2)[0]
    yb = plan.process([None], n_samples=T // 2)[0]
    torch.cuda.synchronize()
    y = torch.cat([ya, yb], dim=1)
    assert bool((y == y[0:1]).all())
    d = np.zeros((1, T), np.float32); d[0, 0] = 1
    ref = _oracle(expr, [d])[0]
    assert np.array_equal(y[C - 1].cpu().numpy(), ref[0])
    assert np.abs(ref).max() > 1.0                       # it really oscillates


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_buffer_input_behind_a_synthesised_one(zg, layout):
    """the buffer wire is input 1, input 0 is synthesised: stage wire k belongs to input k either way"""
    expr = "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])"
    C, T = 45, 500
    x1 = fo.noise(C, T, seed=72)
    d = np.zeros((C, T), np.float32); d[:, 0] = 1
    ys, _ = _run(zg, expr, [d, x1], zg.MODE_EXACT, layout, input_kind=[zg.IN_DIRAC, zg.IN_BUFFER], blocks=[70, 430])
    assert np.array_equal(ys[0], _oracle(expr, [d, x1])[0])


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_six_inputs_two_outputs(zg, layout, dtype):
    """many wires per stage: the geometry has to give up warps per CTA for shared memory"""
    expr = "(_1 + _2 + _3 , _4*_5 - _6[_1])"
    g = zg.compile(expr)
    assert (g.n_in, g.n_out) == (6, 2)
    C, T = 200, 300
    x = [fo.noise(C, T, seed=80 + k) for k in range(6)]
    if dtype == "f32":
        ys, _ = _run(zg, expr, x, zg.MODE_EXACT, layout)
        ref = _oracle(expr, x)
        for y, r in zip(ys, ref):
            assert np.array_equal(y, r)
    else:
        ys, _ = _run_bf16(zg, expr, x, zg.MODE_EXACT, layout)
        ref = _oracle(expr, [fo.bf16_round(v) for v in x])
        for y, r in zip(ys, ref):
            assert np.array_equal(y, fo.bf16_bits(r))


@pytest.mark.parametrize("lanes", [1, 4])
def test_coefficients_computed_on_the_device(zg, lanes):
    """per-voice RBJ low-pass sections computed on the GPU from per-voice cutoff frequencies
    (formulae of reactive_equations/reactive_filter_coeff.cpp:16-50, flowz sign convention) and handed to the plan
    with zg_param_set_device -- no host round trip.  The kernel must use exactly those values: the output equals
    the oracle fed the same coefficients (read back only for the check), bit for bit."""
    torch = _torch()
    S, C, T = 4, 96, 1200
    expr = fo.biquad_cascade_params(S)
    g = zg.compile(expr)
    plan = g.plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=lanes)
    c = torch.arange(C, device="cuda", dtype=torch.float64)
    coefs = []
    for k in range(S):
        f = 440.0 * 2 ** k * (1.0 + c / C)
        w0 = 2.0 * torch.pi * f / 44100.0
        alpha = torch.sin(w0) / (2.0 * 0.707)
        a0 = 1.0 + alpha
        b1 = (1.0 - torch.cos(w0)) / a0
        for v in (b1 / 2.0, b1, b1 / 2.0, 2.0 * torch.cos(w0) / a0, -(1.0 - alpha) / a0):
            coefs.append(v.to(torch.float32).contiguous())
    for i, v in enumerate(coefs):
        plan.set_param_device(i, v)
    x = fo.noise(C, T, seed=91)
    y = plan.process([_to_dev(x)])[0].cpu().numpy()
    assert plan.info().uniform_params == 0 and plan.info().lanes_per_channel == lanes
    params = [v.cpu().numpy() for v in coefs]
    assert np.array_equal(y, _oracle(expr, [x], params)[0])
    # a later host-side set of the same parameter replaces the device-side one
    plan.reset()
    plan.set_param(0, 0.0)                                   # b0 of the first section := 0 for every voice
    params[0] = np.zeros(C, np.float32)
    y2 = plan.process([_to_dev(x)])[0].cpu().numpy()
    assert np.array_equal(y2, _oracle(expr, [x], params)[0])
    with pytest.raises(zg.ZgError):
        plan.set_param_device(0, torch.zeros(C - 1, device="cuda"))


# ---- long delay lines (generated kernel): rings in HBM ------------------------------------------------------
# Delay lines deeper than 16 floats are not register-resident: their state rows are used as a ring, far reads
# become coalesced loads a chunk ahead of use, near reads come from a short register window (zg_ir.hpp).

LONG = [
    "~(_2 + 0.5f*_1[_100])",                                                   # feedback comb
    "_1 + 0.5f*_1[_37] - 0.25f*_1[_1000]",                                     # sparse FIR on the input
    "~(_2 + 0.5f*_1[_3] + 0.25f*_1[_500])",                                    # near and far reads of one line
    "(0.5f*_1 + 0.5f*_1[_1]) |= ~(_2 + 0.7f*_1[_441]) |= (_1 - 0.7f*_1[_20])",  # damped echo, then a notch
    "~(_2 + 0.4f*_1[_17]) |= ~(_2 - 0.3f*_1[_64]) |= (_1 , _1[_33])",            # two long lines, two outputs
]


@pytest.mark.parametrize("expr", LONG)
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_long_delay_lines_exact_and_streaming(zg, expr, layout):
    g = zg.compile(expr)
    C, T = 70, 2600
    x = [fo.noise(C, T, seed=120 + k) for k in range(g.n_in)]
    ref = _oracle(expr, x)
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, layout)
    assert plan.info().jit == 1
    for y, r in zip(ys, ref):
        assert np.array_equal(y, r)
    # ragged blocks, some shorter than the delays: the ring phase follows the stream position
    ys, plan = _run(zg, expr, x, zg.MODE_EXACT, layout, blocks=[1, 7, 31, 32, 33, 100, 5, 1000, 1391])
    for y, r in zip(ys, ref):
        assert np.array_equal(y, r)
    # the state crosses the ABI oldest value first, like the reference's arrays: a fresh plan continues from it
    st = plan.get_state()
    assert st.shape == (g.n_state, C)
    x2 = [fo.noise(C, 700, seed=130 + k) for k in range(g.n_in)]
    full = _oracle(expr, [np.concatenate([a, b], axis=1) for a, b in zip(x, x2)])
    plan2 = g.plan(channels=C, mode=zg.MODE_EXACT, layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED)
    plan2.set_state(st)
    y2 = plan2.process([_to_dev(v.T if layout == "interleaved" else v) for v in x2])
    for y, r in zip(y2, full):
        y = y.cpu().numpy()
        assert np.array_equal(y.T if layout == "interleaved" else y, r[:, T:])


def test_long_delay_lines_state_is_in_reference_order(zg):
    """a pure delay: after T ticks the line holds the last D inputs, oldest first (rotate_push_back)"""
    D, C, T = 300, 40, 1000
    expr = f"_1[_{D}]"
    x = fo.noise(C, T, seed=140)
    ys, plan = _run(zg, expr, [x], zg.MODE_EXACT, blocks=[123, 877])
    want = np.concatenate([np.zeros((C, D), np.float32), x[:, :T - D]], axis=1)
    assert np.array_equal(ys[0], want)
    assert np.array_equal(plan.get_state(), x[:, T - D:].T)
    plan.reset()
    y = plan.process([_to_dev(x[:, :400])])[0].cpu().numpy()
    assert np.array_equal(y, want[:, :400])


def test_long_delay_lines_fast_mode_and_bf16(zg):
    expr = LONG[3]
    C, T = 64, 3000
    x = [fo.noise(C, T, seed=150)]
    ref = _oracle(expr, x)[0]
    ys, _ = _run(zg, expr, x, zg.MODE_FAST)
    assert _rel_err(ys[0], ref) <= TOL
    yb, _ = _run_bf16(zg, expr, x, zg.MODE_EXACT, blocks=[1000, 2000])
    assert np.array_equal(yb[0], fo.bf16_bits(_oracle(expr, [fo.bf16_round(x[0])])[0]))


def test_long_delay_limits_are_reported(zg):
    many = " + ".join(f"0.1f*_1[_{100 + 10 * k}]" for k in range(12))           # 12 far reads of one line: too many
    with pytest.raises(zg.ZgError) as e:
        zg.compile(many).plan(channels=8)
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED and "far reads" in str(e.value)


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_long_delay_lines_through_process_host_chunks(zg, layout):
    """zg_process_host streams a large block in chunks (channel ranges when planar, time ranges when interleaved):
    the rings must follow -- same stream position for every channel chunk, advancing position for time chunks."""
    expr = LONG[3]
    C, T = 4096, 10240                                        # 80 MiB in + 80 MiB out per call: two chunks each
    x = fo.noise(C, T, seed=160)
    plan = zg.compile(expr).plan(channels=C, layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED)
    xin = np.ascontiguousarray(x.T) if layout == "interleaved" else x
    y = plan.process_host([xin[:T // 2] if layout == "interleaved" else np.ascontiguousarray(xin[:, :T // 2])])[0]
    y2 = plan.process_host([xin[T // 2:] if layout == "interleaved" else np.ascontiguousarray(xin[:, T // 2:])])[0]
    assert plan.info().host_chunks > 1
    got = np.concatenate([y, y2], axis=0).T if layout == "interleaved" else np.concatenate([y, y2], axis=1)
    idx = [0, 1, 1023, 1024, 2049, 4095]
    assert np.array_equal(got[idx], _oracle(expr, [x[idx]])[0])


def test_long_delay_lines_set_state_at_any_stream_position(zg):
    """zg_state_set on a plan whose rings are mid-phase (stream position not a multiple of the depths)"""
    expr = LONG[2]
    g = zg.compile(expr)
    C = 33
    xa, xb, xc = (fo.noise(C, n, seed=170 + i) for i, n in enumerate((777, 123, 400)))
    a = g.plan(channels=C)
    a.process([_to_dev(xa)])
    b = g.plan(channels=C)
    b.process([_to_dev(xb)])                                   # b is now at stream position 123, rings mid-phase
    b.set_state(a.get_state())                                 # ... and takes over a's delay lines
    y = b.process([_to_dev(xc)])[0].cpu().numpy()
    ref = _oracle(expr, [np.concatenate([xa, xc], axis=1)])[0][:, 777:]
    assert np.array_equal(y, ref)


def test_small_blocks_replayed_from_a_cuda_graph(zg):
    """Launch-bound use (real-time style 64-sample blocks): zg_process is capturable -- tensor maps travel in the
    kernel parameters, nothing in the call synchronises once the plan is warm -- so a run of block calls can be
    replayed as one CUDA graph.  State lives in device memory, so every replay continues the stream."""
    torch = _torch()
    C, T, NBLK = 4096, 64, 16
    expr = fo.biquad_cascade(4)
    plan = zg.compile(expr).plan(channels=C, mode=zg.MODE_EXACT, lanes_per_channel=1)
    x = fo.noise(C, 2 * NBLK * T, seed=181)
    xin = torch.empty((C, NBLK * T), device="cuda")
    yout = torch.empty_like(xin)
    plan.process([xin[:, :T]], [yout[:, :T]])                 # warm: kernel attributes set, parameters uploaded
    plan.reset()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for b in range(NBLK):
                plan.process([xin[:, b * T:(b + 1) * T]], [yout[:, b * T:(b + 1) * T]])
    plan.reset()                                               # capture ran nothing; start the stream from zero
    outs = []
    for rep in range(2):
        xin.copy_(torch.from_numpy(x[:, rep * NBLK * T:(rep + 1) * NBLK * T]))
        graph.replay()
        torch.cuda.synchronize()
        outs.append(yout.cpu().numpy().copy())
    idx = [0, 17, 4095]
    ref = _oracle(expr, [x[idx]])[0]
    assert np.array_equal(np.concatenate(outs, axis=1)[idx], ref)


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_in_place_blocks(zg, layout):
    """out[j] may be the same buffer as in[i] (zignal_b200.h): `x[t] = f(x[t])` around the reference's tick.  K1, K1b,
    a generated two-wire graph with the wires swapped, bf16 storage; partial overlaps and the FIR kernel refuse."""
    torch = _torch()
    inter = layout == "interleaved"
    C, T = 200, 1000
    lay = zg.INTERLEAVED if inter else zg.PLANAR

    def dev(a):
        return _to_dev(a.T if inter else a)

    def host(t):
        a = t.cpu().numpy()
        return a.T if inter else a

    x = fo.noise(C, T, seed=21)
    expr = fo.biquad_cascade(4)
    want = fo.COracle(expr, C).process([x])[0]
    for lanes in ([1] if inter else [1, 4]):                         # K1, K1b (planar only)
        plan = zg.compile(expr).plan(channels=C, layout=lay, lanes_per_channel=lanes)
        buf = dev(x)
        for t0, n in ((0, 600), (600, T - 600)):                     # two blocks, each overwritten in place
            blk = buf[t0:t0 + n] if inter else buf[:, t0:t0 + n]
            out = plan.process([blk], [blk], n_samples=n)[0]
            assert out.data_ptr() == blk.data_ptr()
        torch.cuda.synchronize()
        assert np.array_equal(host(buf), want), f"lanes={lanes}"

    # generated kernel, two wires, swapped: out[0] overwrites in[1] and out[1] overwrites in[0]
    g2 = "(_1 + _2 , _1 - _2[_1]) |= (~(_2 + 0.5f*_1[_1]) | (_1 - 0.25f*_1[_2]))"
    x2 = [fo.noise(C, T, seed=22), fo.noise(C, T, seed=23)]
    w2 = fo.COracle(g2, C).process(x2)
    a, b = dev(x2[0]), dev(x2[1])
    zg.compile(g2).plan(channels=C, layout=lay).process([a, b], [b, a])
    torch.cuda.synchronize()
    assert np.array_equal(host(b), w2[0]) and np.array_equal(host(a), w2[1])

    # bf16 storage in place
    xb = fo.bf16_round(x)
    wb = fo.COracle(expr, C).process([xb])[0]
    bb = zg.to_block(xb.T if inter else xb, dtype=torch.bfloat16)
    zg.compile(expr).plan(channels=C, layout=lay, io_dtype=zg.BF16).process([bb], [bb])
    torch.cuda.synchronize()
    assert np.array_equal(host(bb.view(torch.int16)).view(np.uint16), fo.bf16_bits(wb))

    # anything else that overlaps is refused before a kernel is launched
    plan = zg.compile(expr).plan(channels=C, layout=lay)
    big = _to_dev(np.zeros((C + 8, T + 8) if not inter else (T + 8, C + 8), np.float32))
    src, dst = (big[:C, :T], big[4:C + 4, :T]) if not inter else (big[:T, :C], big[4:T + 4, :C])
    with pytest.raises(zg.ZgError) as e:
        plan.process([src], [dst])
    assert e.value.status == zg.ZG_ERR_ARG
    fir = zg.compile(fo.fir_expr(fo.fir_taps(32))).plan(channels=C, layout=lay)
    buf = dev(x)
    with pytest.raises(zg.ZgError) as e:
        fir.process([buf], [buf])
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


def test_plans_of_one_prebuilt_kernel_with_different_geometry_alternate(zg):
    """The dynamic shared-memory limit belongs to the kernel function, which all plans of a process share: a small plan
    launched after a large one must not lower it for the large one's next launch (biquad and FIR kernels)."""
    torch = _torch()
    expr = fo.biquad_cascade(4)
    g = zg.compile(expr)
    big_x, small_x = fo.noise(16384, 512, seed=1), fo.noise(33, 64, seed=2)
    big = g.plan(channels=16384, lanes_per_channel=1)
    small = g.plan(channels=33, lanes_per_channel=1)
    want_big, want_small = _oracle(expr, [big_x])[0], _oracle(expr, [small_x])[0]
    for _ in range(3):
        big.reset(); small.reset()
        yb = big.process([_to_dev(big_x)])[0]
        ys = small.process([_to_dev(small_x)])[0]
        torch.cuda.synchronize()
        assert big.info().smem_bytes > small.info().smem_bytes
        assert np.array_equal(yb.cpu().numpy(), want_big) and np.array_equal(ys.cpu().numpy(), want_small)
    h256, h32 = fo.fir_taps(256), fo.fir_taps(32)
    f_big, f_small = zg.compile(fo.fir_expr(h256)).plan(channels=64), zg.compile(fo.fir_expr(h32)).plan(channels=64)
    xf = fo.noise(64, 512, seed=3)
    for _ in range(3):
        f_big.reset(); f_small.reset()
        a = f_big.process([_to_dev(xf)])[0].cpu().numpy()
        b = f_small.process([_to_dev(xf)])[0].cpu().numpy()
        assert np.array_equal(a, fo.fir_direct(xf, h256)) and np.array_equal(b, fo.fir_direct(xf, h32))
