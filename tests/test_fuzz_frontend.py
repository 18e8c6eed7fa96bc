"""Random flowz expressions: the product's front end and host voice against the oracle (CPU only).

Every expression is generated from the grammar (flowz.hpp:68-102); for each one the two sides must agree on whether it
is a valid graph at all, on arity, per-wire max/min delays, the canonical form, the ResultType of the bare expression,
and on six ticks of one voice with small-integer inputs (exact in int, float and double alike).  A third evaluator of different
construction (tests/netlist_flowz.py: a netlist of the uncanonicalised expression, evaluated by need) ticks along; it
alone checks the graphs the product accepts beyond the reference -- feedbacks that cannot be split into a promise and
a future part but still have a delay in every loop (nested feedback, TODO.md:11-27).  The seeds are fixed: the test is
deterministic."""
import random

import pytest

import flowz_oracle as fo
import netlist_flowz as nl

CONSTS = ["2", "3", "0.5f", "0.25f", "1.5", "-1", "0x1p-1f", "-0.75f", "$0", "$1"]
PARAMS = [0.5, -0.25]                       # values of the std::ref terminals $0, $1


def _gen(rng, depth, wires, scalar=False, consts=CONSTS, ops="+-*+-*/"):
    """A random expression over placeholders _1.._wires; scalar=True: leaf arithmetic only (what an operand of + - *
    must be; a combinator there is ill-formed in the reference, and one in ten operands is generated that way)."""
    if depth <= 0 or rng.random() < 0.2:
        k = rng.randint(1, wires)
        r = rng.random()
        if r < 0.45:
            return f"_{k}"
        if r < 0.8:
            return f"_{k}[_{rng.randint(1, 3)}]"
        return rng.choice(consts)
    r = rng.random() * (0.45 if scalar else 1.0)
    sub_scalar = r < 0.45 and rng.random() < 0.9
    a, b = _gen(rng, depth - 1, wires, sub_scalar, consts, ops), _gen(rng, depth - 1, wires, sub_scalar, consts, ops)
    if r < 0.40:
        return f"({a} {rng.choice(ops)} {b})"                          # by default one operator in seven is a division
    if r < 0.45:
        return f"(-{a})"
    if r < 0.62:
        return f"({a} |= {b})"
    if r < 0.72:
        return f"({a} | {b})"
    if r < 0.84:
        return f"({a} , {b})"
    return f"(~{a})"


def _same(a, b):
    """Tuples of floats, NaN equal to NaN (x / 0 and 0 / 0 are inf and NaN in float and double; 0 in int on both sides)."""
    import math
    return len(a) == len(b) and all(x == y or (math.isnan(x) and math.isnan(y)) for x, y in zip(a, b))


def _product(zg, expr):
    try:
        g = zg.compile(expr)
    except zg.ZgError as e:
        return None, ("limit" if e.status == zg.ZG_ERR_UNSUPPORTED else str(e))
    return g, None


def _voice(g):
    v = g.voice()
    for k in range(g.n_params):
        v.set_param(k, PARAMS[k])
    return v


def _oracle(expr):
    try:
        o = fo.Oracle(expr, params=PARAMS)
        n_in = fo.input_arity(fo.parse(expr))
        res = fo.Oracle(expr, params=PARAMS).tick(*([0.0] * n_in))   # ill-formed operands only surface when the walk reaches them
        if any(r is fo.BOTTOM for r in res):          # the reference would hand back a bottom_type (flowz.hpp:1004)
            return None, "an output is a fed-back wire nothing ever assigns"
        return o, None
    except Exception as e:                      # the restatement signals invalid graphs with plain exceptions
        return None, str(e)


@pytest.mark.parametrize("seed", range(16))
def test_random_graphs_product_equals_oracle(zg, seed):
    rng = random.Random(1000 + seed)
    checked = rejected = extended = 0
    for _ in range(60):
        expr = _gen(rng, rng.randint(1, 5), rng.randint(1, 4))
        g, perr = _product(zg, expr)
        o, oerr = _oracle(expr)
        if perr == "limit":                          # more than ZG_MAX_WIRES wires: a limit of the product, not of flowz
            continue
        if g is not None and o is None:
            # beyond the reference: a feedback that cannot be split, or whose split would make the reference touch the
            # current value of a fed-back wire (its bottom_type) -- the product resolves fed-back wires as forward
            # references and accepts whatever has a delay in every loop; the netlist evaluator is the only other opinion
            assert any(m in oerr for m in ("cannot be split", "fed-back wire", "not a single value")), f"{expr}: {oerr}"
            net = nl.Netlist(expr, params=PARAMS)
            dt = [rng.choice([fo.I32, fo.F32, fo.F64]) for _ in range(g.n_in)]
            v = _voice(g)
            for t in range(6):
                xs = [float(rng.randint(-3, 3)) for _ in range(g.n_in)]
                res = net.tick(*xs, dtype=dt)
                assert _same(tuple(float(y) for y in v.tick(*xs, dtypes=dt)), tuple(float(x) for _, x in res)), expr
                assert v.out_dtypes == tuple(d for d, _ in res), expr
            extended += 1
            continue
        assert (g is None) == (o is None), f"{expr}\n product: {perr}\n oracle: {oerr}"
        if g is None:
            rejected += 1
            continue
        e = fo.parse(expr)
        n_in, n_out = fo.input_arity(e), fo.output_arity(e)
        assert zg.arity(expr) == (n_in, n_out), expr
        n_tick = len(fo.Oracle(expr, params=PARAMS).tick(*([0.0] * n_in)))         # what one tick returns (flowz.hpp:996-999 can exceed
        assert g.n_out == n_tick, (expr, g.n_out, n_tick)           # output_arity: surplus inputs pass through a sequence)
        assert zg.delays(expr) == fo.max_input_delays(e), expr
        assert zg.delays(expr, minimum=True) == fo.min_input_delays(e), expr
        c = zg.canonical(expr)
        assert c == zg.canonical(str(fo.make_canonical(e))), expr
        assert zg.canonical(c) == c, expr                                     # printed trees read back as themselves
        try:
            want_t = fo.result_type(e, [fo.F32] * n_in)
        except TypeError:
            with pytest.raises(zg.ZgError):
                zg.result_types(expr, [zg.F32] * n_in)
        else:
            assert zg.result_types(expr, [zg.F32] * n_in) == want_t, expr
        v = _voice(g)
        net = nl.Netlist(expr, params=PARAMS) if "~" not in expr else None
        dt = [rng.choice([fo.I32, fo.F32, fo.F64]) for _ in range(n_in)]   # the C++ type of each argument: int stays int (:769-772)
        for t in range(6):
            xs = [float(rng.randint(-3, 3)) for _ in range(n_in)]
            res = o.tick(*xs, dtype=dt)
            want = tuple(float(val[0]) for _, val in res)
            if net is not None:      # (inside a feedback the reference's split can route external inputs differently from
                #  the plain reading of the expression -- its "thinning" bug, TODO.md:11 -- and product and oracle follow it)
                third = net.tick(*xs, dtype=dt)
                assert [d for d, _ in third] == [d for d, _ in res] and _same([float(x) for _, x in third], want), expr
            got = tuple(float(y) for y in v.tick(*xs, dtypes=dt))
            assert _same(got, want), f"{expr} tick {t}: {got} != {want}"
            assert v.out_dtypes == tuple(d for d, _ in res), f"{expr}: result types {v.out_dtypes} != {[d for d, _ in res]}"
        checked += 1
    assert checked >= 10, (checked, rejected)
