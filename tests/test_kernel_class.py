"""Host-side recognisers (zg_graph_kernel_class, zg_match.cpp): which device kernel family a tick
program is routed to.  A graph is only routed to a prebuilt kernel when its lowered arithmetic is
exactly the kernel's, association included; everything else must fall through to the generated
kernel.  CPU only.  Also pins the vectorised FIR oracle against the tick-by-tick oracles."""
import numpy as np
import pytest

import flowz_oracle as fo


@pytest.mark.parametrize("expr,want", [
    (fo.fir_expr(fo.fir_taps(256)), "fir:256"),
    (fo.fir_expr(fo.fir_taps(2)), "fir:2"),
    (fo.fir_expr_params(7), "fir:7"),
    ("_1[_1]*0.5f + 0.25f*_1", "generated"),                          # taps out of order
    ("0.5f*_1 + 0.25f*_1[_2]", "generated"),                          # sparse
    ("0.5f*_1 + (0.25f*_1[_1] + 0.125f*_1[_2])", "generated"),        # right-associated sub-sum
    ("0.5f*_1 + 0.25f*_1[_1] - 0.125f*_1[_2]", "generated"),          # a difference is not a sum
    ("0.5f*_1", "generated"),
    (fo.biquad_cascade(4), "biquad_df1:4"),
    (fo.biquad_cascade_params(2), "biquad_df1:2"),
    (fo.biquad_cascade(9), "generated"),                              # more sections than prebuilt
    ("~(_2 + 0.9f*_1[_1])", "generated"),
    ("_1 + 1", "host-only"),
    ("0.5*_1", "host-only"),
])
def test_kernel_class(zg, expr, want):
    assert zg.compile(expr).kernel_class() == want


def test_fir_accepts_both_delay_spellings_and_operand_orders(zg):
    assert zg.compile("0.5f*_1 + _1[-1]*0.25f + 0.125f*_1[_2]").kernel_class() == "fir:3"
    assert zg.compile("_1*0.5f >> 1.0f*_1 + 0.25f*_1[-1]").kernel_class() == "generated"


@pytest.mark.parametrize("n", [2, 5, 16, 33, 256])
def test_vectorised_fir_oracle_equals_tick_oracles(zg, n):
    """fir_direct (used for the large GPU parity cases) == the emitted-C tick oracle == the numpy tick
    oracle == the library's own host voice, bit for bit."""
    h = fo.fir_taps(n)
    expr = fo.fir_expr(h)
    x = fo.noise(3, 300, seed=n)
    want = fo.fir_direct(x, h)
    assert np.array_equal(fo.COracle(expr, 3).process([x])[0], want)
    if n <= 33:
        assert np.array_equal(fo.Oracle(expr, 3).process([x])[0], want)
    v = zg.compile(expr).voice()
    got = np.array([v.tick(np.float32(s))[0] for s in x[1, :64]], np.float32)
    assert np.array_equal(got, want[1, :64])
    # with history: continuing a stream
    assert np.array_equal(fo.fir_direct(x[:, 100:], h, history=np.concatenate(
        [np.zeros((3, max(0, n - 1 - 100)), np.float32), x[:, max(0, 100 - (n - 1)):100]], axis=1)), want[:, 100:])


def test_package_workloads_equal_the_oracles_copies(zg):
    """bench.py's GPU arm and the tools build their graphs from zignal_b200/workloads.py (nothing under
    oracle/ is imported there); the oracle keeps its own builders.  Same text, same coefficients."""
    from zignal_b200 import workloads as wl
    for n in (1, 2, 4, 8):
        assert wl.biquad_cascade(n) == fo.biquad_cascade(n)
        assert wl.biquad_cascade_params(n) == fo.biquad_cascade_params(n)
    assert wl.osc_lp_expr() == fo.osc_lp_expr() and wl.poly_voice_expr() == fo.poly_voice_expr()
    for n in (2, 17, 256, 512):
        assert np.array_equal(np.array(wl.fir_taps(n), np.float32), fo.fir_taps(n))
        assert wl.fir_expr(wl.fir_taps(n)) == fo.fir_expr(fo.fir_taps(n))
    for f in (440.0, 3520.0, 440.0 * 1.37):
        assert wl.rbj_lowpass(f) == tuple(float(v) for v in fo.rbj_lowpass(f))


def test_linearity_of_the_tick_program(zg):
    import flowz_oracle as fo
    L, A, N = zg.LINEAR, zg.AFFINE, zg.NONLINEAR
    cases = [(fo.biquad_cascade(4), L), (fo.biquad_cascade_params(2), L), (fo.osc_lp_expr(), L), (fo.poly_voice_expr(), L),
             (fo.fir_expr(fo.fir_taps(16)), L), ("~(_2 + 0.5f*_1[_441]) |= (_1 + 0.25f*_1[_1000])", L), ("_1/2", L),
             ("~(_2 + $0*_1[_1])", L), ("_1 + 1", A), ("0*_1 + 3", A), ("~(_2 + 0.5f*_1[_1] + 0.125f)", A),
             ("_1*_1", N), ("_1*_1[_1]", N), ("_1/_2", N), ("~(_2 + _1[_1]*_1[_2])", N)]
    for expr, want in cases:
        assert zg.compile(expr).linearity() == want, expr


def test_linear_graphs_obey_superposition(zg):
    """Property behind the label: for every random graph reported LINEAR, tick(a + b) == tick(a) + tick(b) with the
    state carried along, exactly (small integers and dyadic constants: no rounding anywhere); AFFINE graphs obey it
    after subtracting the zero-input response; and some NONLINEAR graph must break it."""
    import random
    from test_fuzz_frontend import _gen
    rng = random.Random(77)
    consts = ["2", "3", "0.5f", "0.25f", "-1", "0x1p-1f", "-0.75f"]
    seen = {zg.LINEAR: 0, zg.AFFINE: 0, zg.NONLINEAR: 0}
    broke = 0
    for _ in range(400):
        e = _gen(rng, rng.randint(1, 4), rng.randint(1, 3), consts=consts, ops="+-*")
        try:
            g = zg.compile(e)
        except zg.ZgError:
            continue
        if g.n_in == 0:
            continue
        kind = g.linearity()
        seen[kind] += 1
        va, vb, vs, v0 = g.voice(), g.voice(), g.voice(), g.voice()
        ok = True
        for t in range(5):
            a = [float(rng.randint(-3, 3)) for _ in range(g.n_in)]
            b = [float(rng.randint(-3, 3)) for _ in range(g.n_in)]
            ya, yb, ys, y0 = va(*a), vb(*b), vs(*[p + q for p, q in zip(a, b)]), v0(*([0.0] * g.n_in))
            ok = ok and all(s - z == (p - z) + (q - z) for s, p, q, z in zip(ys, ya, yb, y0))
            if kind == zg.LINEAR:
                assert all(z == 0 for z in y0), e
        if kind != zg.NONLINEAR:
            assert ok, e
        else:
            broke += not ok
    assert min(seen.values()) >= 10 and broke >= 5, (seen, broke)


def test_state_space_of_linear_graphs(zg):
    """state' = A state + B x, y = C state + D x read off the tick program reproduces the ticks (float64 recursion against
    the fp32 host voice: agreement to what fp32 rounding noise allows on these recursions -- the oscillator is marginally
    stable, the 440 Hz section has a noise gain of ~40), and the poles of the benchmark biquad are where the RBJ design puts them."""
    import flowz_oracle as fo
    rng = np.random.default_rng(5)
    for expr, params in [(fo.biquad_cascade(4), ()), (fo.osc_lp_expr(), ()), ("~(_2 + $0*_1[_1]) |= _1 - 0.5f*_1[_3]", (0.75,)),
                         ("(_1 + _2 , _1 - _2[_1]) |= (~(_2 + 0.5f*_1[_1]) | (_1 - 0.25f*_1[_2]))", ())]:
        g = zg.compile(expr)
        A, B, Cm, D = g.state_space(params)
        assert A.shape == (g.n_state, g.n_state) and D.shape == (g.n_out, g.n_in)
        v = g.voice()
        for k, p in enumerate(params):
            v.set_param(k, p)
        x, peak = np.zeros(g.n_state), 0.0
        for t in range(200):
            u = rng.uniform(-1, 1, g.n_in).astype(np.float32)
            y = np.array(v.tick(*[float(a) for a in u], dtypes=[zg.F32] * g.n_in))
            want = Cm @ x + D @ u
            peak = max(peak, float(np.abs(want).max()))
            assert np.allclose(y, want, rtol=0, atol=2e-3 * max(peak, 1.0)), (expr, t, y, want)
            x = A @ x + B @ u
    # 4 x RBJ low-pass: every section contributes a conjugate pole pair inside the unit circle (plus the zeros of the shift rows)
    A = zg.compile(fo.biquad_cascade(4)).state_space()[0]
    poles = np.linalg.eigvals(A)
    assert np.abs(poles).max() < 1.0 and (np.abs(poles) > 0.5).sum() == 8
    with pytest.raises(ValueError):
        zg.compile("_1*_1").state_space()
