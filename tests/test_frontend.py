"""The product's host side (parser, static analysis, canonical split, lowering, host voice) against
the reference's own vectors and against the oracle.  CPU only."""
import numpy as np
import pytest

import flowz_oracle as fo
import reference_vectors as rv


def _norm(zg, text):
    # canonical() is the identity on trees without '~': use it as parse -> print normalisation
    return zg.canonical(text)


@pytest.mark.parametrize("expr,left,right,line", rv.CANONICAL_SPLITS, ids=[f"tests.cpp:{c[3]}" for c in rv.CANONICAL_SPLITS])
def test_canonical_split(zg, expr, left, right, line):
    assert zg.canonical(expr) == _norm(zg, f"bfb({left} , {right})")
    # and the product's printer agrees with the oracle's on the same tree
    assert zg.canonical(expr) == _norm(zg, str(fo.make_canonical(fo.parse(expr))))


@pytest.mark.parametrize("expr,n_in,n_out,delays,line", rv.CANONICAL_ARITY)
def test_canonical_arity(zg, expr, n_in, n_out, delays, line):
    c = zg.canonical(expr)
    assert zg.arity(c) == (n_in, n_out)
    assert zg.delays(c) == delays


@pytest.mark.parametrize("expr,n_in,n_out,ins,outs,line", rv.WIRES_AROUND)
def test_wires_around_boxes(zg, expr, n_in, n_out, ins, outs, line):
    assert zg.arity(expr) == (n_in, n_out)
    assert zg.compile(expr).voice()(*ins) == outs


@pytest.mark.parametrize("expr,steps,line", rv.TICKS, ids=[f"{c[0]}@{c[2]}" for c in rv.TICKS])
def test_known_answer_ticks(zg, expr, steps, line):
    v = zg.compile(expr).voice()
    for ins, outs in steps:
        got = v(*ins)
        assert got == outs
        assert all(isinstance(g, int) for g in got) or "[" in expr   # ints stay ints unless they went through a float line


@pytest.mark.parametrize("expr,types,is_tuple,canonical,line", rv.RESULT_TYPES, ids=[f"tests.cpp:{c[4]}" for c in rv.RESULT_TYPES])
def test_result_type(zg, expr, types, is_tuple, canonical, line):
    # ResultType, flowz.hpp:515-644; input tuple<float> as in test/tests.cpp:192
    text = zg.canonical(expr) if canonical else expr
    assert zg.result_types(text, [zg.F32]) == ([rv.TYPE_CODE[c] for c in types], is_tuple)


def test_result_type_agrees_with_the_oracle_beyond_the_reference_vectors(zg):
    # the product resolves absorbers in one pass, the oracle keeps the reference's two passes (:594-609)
    import itertools
    exprs = ["~(_1[_1] + _2 + _3)", "~(_1[_1]*2 + _2) |= _1*_1", "(_1 | _1*1.0) |= (_2,_1,_1+_2)", "~(_1[_1])",
             "~(-_1[_1] + _2)", "~( (_1[_1] + _2)*0.5f |= _1 )", "_1*$0 |= ~(_2 + $1*_1[_1])", "~~( _1[_1] + _2[_1] )",
             "(_1, _2) |= ~(_1[_1] + _2 + _3)", "_1/2 |= _1*_1[_1]", "~(_1[_1] + _2*cplx{0,1})", "_1 | ~(_1[_2]*_2)"]
    for expr in exprs:
        n_in = zg.arity(expr)[0]
        for sig in itertools.product([zg.I32, zg.F32, zg.F64], repeat=n_in):
            try:
                want = fo.result_type(expr, list(sig))
            except TypeError:
                with pytest.raises(zg.ZgError):
                    zg.result_types(expr, list(sig))
                continue
            assert zg.result_types(expr, list(sig)) == want, (expr, sig)
    assert zg.result_types("~(_1[_1])", [zg.F32]) == ([zg.TYPE_OPEN], True)      # the reference's leftover absorber (:575-578)
    with pytest.raises(zg.ZgError):                                              # complex<float> * double: no such operator
        zg.result_types("_1*cplx{1,0}", [zg.F64])
    with pytest.raises(zg.ZgError):                                              # typed, but not evaluated
        zg.compile("_1*cplx{1,0}")


def test_spellings_the_reference_plans(zg):
    # TODO.md:8-9  `_1<-2>` = input one delayed by 2;  TODO.md:51-52  `(_1+_2)[_1]` is equivalent to `_1+_2 |= _1[_1]`
    assert zg.canonical("_1<-2> + _2") == zg.canonical("_1[_2] + _2")
    assert zg.canonical("(_1+_2)[_1]") == zg.canonical("_1+_2 |= _1[_1]")
    assert zg.canonical("(_1 , 2*_1)[_3]") == zg.canonical("(_1 , 2*_1) |= (_1[_3] | _1[_3])")
    assert zg.canonical("_1[_1][_2]") == zg.canonical("_1[_1] |= _1[_2]")
    a, b = zg.compile("~((_1 + _2)[_1])").voice(), fo.Oracle("~(_1 + _2 |= _1[_1])")
    for t in range(6):
        assert a(float(t + 1)) == tuple(float(v[0]) for _, v in b.tick(float(t + 1)))
    with pytest.raises(zg.ZgError):
        zg.canonical("_1<2>")


def test_feedbacks_the_reference_cannot_compile(zg):
    """Nested feedback (TODO.md:11-27; the disabled expectation test/tests.cpp:59) and parallel combiners inside a loop
    (TODO.md:29): the reference cannot split these, or its split touches the current value of a fed-back wire.  compile()
    takes such graphs at their word -- `~x` ties the first inputs of x to x's outputs -- and refuses only loops without
    a delay.  Expected sequences derived by hand; tests/netlist_flowz.py agrees."""
    import netlist_flowz as nl
    # s[t] = y + y + 1 with y = s[t-1]:  y = 0, 1, 3, 7, 15, 31
    g = zg.compile("~~( _1 + _2 + 1 |= _1[_1] )")
    v = g.voice()
    assert [v()[0] for _ in range(6)] == [0.0, 1.0, 3.0, 7.0, 15.0, 31.0]
    assert (g.n_in, g.n_out, g.n_state) == (0, 1, 1)
    with pytest.raises(zg.ZgError):                                  # the reference-shaped analysis still says no
        zg.canonical("~~( _1 + _2 |= _1[_1] )")
    # the TODO's picture: f = 0.5*, inner loop integrates: y[t] = s[t-1], s[t] = y + (0.5 y + u)  ->  impulse response 1.5^t
    v = zg.compile("~( (0.5f*_1 + _2) |= ~(_1 + _2 |= _1[_1]) )").voice()
    assert [v(u)[0] for u in (1.0, 0.0, 0.0, 0.0, 0.0)] == [0.0, 1.0, 1.5, 2.25, 3.375]
    # parallel combiner inside the loop
    e = "~((_1[_2] |= _2) | (_1 - _2))"
    g, net = zg.compile(e), nl.Netlist(e)
    assert (g.n_in, g.n_out) == (2, 2)
    v = g.voice()
    for t in range(8):
        assert v(float(t + 1), 2.0) == tuple(float(x) for _, x in net.tick(float(t + 1), 2.0))
    # a graph the reference does compile keeps the reference's walk, state layout included
    assert zg.compile("~(_2 + 0.5f*_1[_1])").canonical == zg.canonical("_1 |= ~(_2 + 0.5f*_1[_1])")
    # loops without a delay stay errors, with a message that says so
    for bad in ["~(_1 + _2)", "~(_1)", "~(2*_1)", "~~(_1 + _2 |= _1)"]:
        with pytest.raises(zg.ZgError, match="without a delay"):
            zg.compile(bad)
    # undefined behaviour in the reference is an error, not a silent read of a neighbouring line (flowz.hpp:950-958, 1043-1047)
    with pytest.raises(zg.ZgError, match="out-of-bounds in the reference"):
        zg.compile("~(((_3 * _1) + (_2 - _3[_1])) |= ((1.5 - _2) * (_3[_2] + _3)))")
    with pytest.raises(ValueError, match="out-of-bounds in the reference"):
        fo.Oracle("~(((_3 * _1) + (_2 - _3[_1])) |= ((1.5 - _2) * (_3[_2] + _3)))").tick(1.0, 2.0, 3.0, 4.0)


def test_callable_arity_is_the_user_expressions(zg):
    # compile() takes arity_t from the expression as written (flowz.hpp:1238); the canonical tree may count more inputs
    # (the split can move a sub-expression with unused inputs into the promise part) without anything reading them.
    # Found by fuzzing: the product used to insist on the canonical tree's count.
    for e in ["~((_1 |= _3) , (_2 / _2[_2]))", "~((_1 |= _2) * (((_1[_3] * _2) - (_2[_2] * _2)) * 2))",
              "~((_1 | _2) |= (((_1[_1] |= $0) |= (_2 |= _1[_2])) |= _1))"]:
        g, o = zg.compile(e), fo.Oracle(e, params=[0.5])
        assert g.n_in == zg.arity(e)[0] == fo.input_arity(fo.parse(e))
        v = g.voice()
        if g.n_params:
            v.set_param(0, 0.5)
        for t in range(6):
            xs = [float(t + 1 + k) for k in range(g.n_in)]
            assert v(*xs) == tuple(float(x[0]) for _, x in o.tick(*xs))


def test_expression_depth_is_bounded_not_the_stack(zg):
    # every analysis recurses over the tree: trees up to 1536 levels are accepted (a 512-tap FIR sum is 513 high), deeper
    # text is an error, never a stack overflow; whatever is accepted prints to text that parses back to the same tree
    for e in ["(" * 1500 + "_1" + ")" * 1500, "-" * 1500 + "_1", "_1" + " + 0.5f*_1[_1]" * 1500, "_1" + " |= _1" * 1500]:
        c = zg.canonical(e)
        assert zg.canonical(c) == c
        assert zg.arity(c) == zg.arity(e) == (1, 1)
        zg.compile(e).voice()(1.0)
    for e in ["(" * 5000 + "_1" + ")" * 5000, "-" * 5000 + "_1", "~" * 5000 + "_1", "_1" + " + _1" * 5000, "_1" + " |= _1" * 5000,
              "_1" + "[_1]" * 3000]:
        with pytest.raises(zg.ZgError) as err:
            zg.arity(e)
        assert err.value.status in (zg.ZG_ERR_PARSE, zg.ZG_ERR_GRAPH) and len(str(err.value)) < 400


def test_series_and_delay_spellings_of_the_prototypes(zg):
    # north star: `>>` and `_1[-n]` (experimental_steps/wires_mono_only.cpp:37, delay_expression.cpp:99)
    assert zg.canonical("_1 >> _1[-1]") == zg.canonical("_1 |= _1[_1]")
    assert zg.canonical("~(_2 + 0.5f*_1[-1]) >> _1[-2]") == zg.canonical("~(_2 + 0.5f*_1[_1]) |= _1[_2]")
    # >> is left associative and binds tighter than | and |=
    assert zg.canonical("_1 >> _1 >> _1") == zg.canonical("(_1 |= _1) |= _1")


def test_clone_copies_state(zg):
    v = zg.compile("~(_1[_1] + _2)").voice()
    v(5)
    w = v.clone()                          # flowz.hpp:1206-1207: copying the callable copies state_
    assert v(1) == (6,) and w(10) == (15,)


def test_std_ref_parameter_modulation(zg):
    # flowz/README.md:42-63: ~( std::ref(a)*_1[_1] + 0.1*_2 ), a *= 0.9f per sample; 0.1 is a double
    g = zg.compile("~($0*_1[_1] + 0.1*_2)")
    v = g.voice()
    o = fo.Oracle("~($0*_1[_1] + 0.1*_2)", params=[0.0])
    a = np.float32(0.9)
    for t in range(10):
        x = 1.0 if t == 0 else 0.0
        v.set_param(0, float(a))
        o.backend.params[0] = a
        (dt, want), = o.tick(x)
        got, = v(x)
        assert dt == fo.F64 and got == float(want[0])
        a = np.float32(a * np.float32(0.9))


GRAPHS = [
    "~(_2 + 0.9f*_1[_1])",                                             # BASELINE config 1
    "~(0x1.fcp0f*_1[_1] - _1[_2] + _2) |= ~(_2 + 0.9f*_1[_1])",        # config 3: osc >> one-pole
    "_1 |= (_1[_1] , _1[_3]) |= _1 - _2",
    "(_1 , _1[_2]) |= (_1 | 0.5f*_1) |= _1*_2",
    "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])",
    "~( (_2 + 0.3f*_1[_1]) |= (0.5f*_1 + 0.25f*_1[_1]) )",
    "_1 / (_2*_2 + 1.5f) - -_1[_1]",
]


@pytest.mark.parametrize("expr", GRAPHS + [rv.bench_graphs()[k] for k in (1, 2, 3, 4)] + [fo.biquad_cascade(4)])
def test_host_voice_equals_oracle(zg, expr):
    g = zg.compile(expr)
    o = fo.Oracle(expr)
    assert (g.n_in, g.n_out) == (o.n_in, o.n_out)
    assert g.canonical == zg.canonical(str(o.canonical))     # same tree (printer-normalised)
    x = fo.noise(g.n_in, 400, seed=11)
    v = g.voice()
    for t in range(400):
        want = o.tick(*[x[i, t] for i in range(g.n_in)])
        got = v(*[float(x[i, t]) for i in range(g.n_in)])
        for (dt, w), y in zip(want, got):
            assert np.float32(y) == w[0] or (np.isnan(y) and np.isnan(w[0])), (expr, t)


def test_line_sharing_removes_duplicate_state(zg):
    # DF2: the reference keeps two copies of the u line; the tick program keeps one
    g = zg.compile(rv.bench_graphs()[2])
    assert g.n_state == 2
    g4 = zg.compile(fo.biquad_cascade(4))
    assert (g4.n_state, g4.n_lines) == (10, 5)


@pytest.mark.parametrize("expr", ["_1 +", "_0", "_1[_0]", "~(_1 + _2)", "(_1 |= _1) + _1", "foo", "_1 |= |= _1"])
def test_errors_are_reported_not_thrown(zg, expr):
    with pytest.raises(zg.ZgError) as e:
        zg.compile(expr)
    assert e.value.status in (zg.ZG_ERR_PARSE, zg.ZG_ERR_GRAPH)


def test_mono_one_pole_1024_samples_on_cpu(zg):
    """BASELINE config 0: y = x + a*_1[-1] over 1024 samples, compile() plumbing without a GPU."""
    a = np.float32(0.9)
    v = zg.compile("~(_2 + 0.9f*_1[-1])").voice()
    x = fo.noise(1, 1024, seed=2)[0]
    y1 = np.float32(0)
    for t in range(1024):
        y1 = np.float32(x[t] + np.float32(a * y1))
        assert np.float32(v(float(x[t]))[0]) == y1
