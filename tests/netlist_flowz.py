"""A third, structurally different evaluator for flowz graphs -- TEST INFRASTRUCTURE.

The product lowers the canonical tree to SSA with delay lines; the oracle (oracle/flowz_oracle.py) walks the canonical
tree with a nested state tree like the reference.  This one never canonicalises: it turns the *user's* expression into
a netlist of wires once (placeholders are just references to wires, a feedback `~x` ties the first inputs of x to its
own outputs through forward cells), and evaluates every tick by need, with a history per wire for delayed reads.  A
wire that is asked for its value while it is being computed is a loop without a delay.

It therefore defines flowz semantics as the fixed point the reference's evaluators compute in a fixed order
(sequence flowz.hpp:960-1001 incl. the pass-through of surplus inputs :996-999, parallel :1076-1101, channel :765-768,
binary_feedback :1031-1074, state narrowed to float :136, C++ arithmetic on the argument types :769-772), and extends
to the nested feedbacks the reference cannot split (TODO.md:11-27, disabled test/tests.cpp:59).  Used by
tests/test_fuzz_frontend.py for graphs where the oracle has no opinion, and as a third opinion where it has one."""
import numpy as np

import flowz_oracle as fo

_NP = {fo.I32: np.int32, fo.F32: np.float32, fo.F64: np.float64}


class Loop(Exception):
    pass


class Wire:
    """kind: in | const | param | op | neg | delay | cell"""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)
        self.hist = []              # values of past ticks, narrowed to float (only kept for delay sources)
        self.busy = False


class Netlist:
    def __init__(self, expr, params=()):
        self.tree = fo.parse(expr) if isinstance(expr, str) else expr
        self.n_in = fo.input_arity(self.tree)
        self.params = list(params)
        self.inputs = [Wire("in", k=k) for k in range(self.n_in)]
        self.delay_sources = []
        self.outputs = self._walk(self.tree, list(self.inputs), list(self.inputs))
        self.memo = {}

    # ---- netlist construction: routing only, nothing is evaluated --------------------------------------------
    def _delay(self, src, n):
        if src not in self.delay_sources:
            self.delay_sources.append(src)
        return Wire("delay", src=src, n=n)

    def _scalar(self, e, ins, dins):
        out = self._walk(e, ins, dins)
        if len(out) != 1:
            raise ValueError("arithmetic on a multi-wire sub-expression")
        return out[0]

    def _walk(self, e, ins, dins):
        """ins: the wires whose CURRENT values the placeholders of e see; dins: the wires whose HISTORY `_k[_n]` reads.
        The two lists name the same wires except where the reference lets them drift apart: a sequence hands its right
        side the left side's actual results plus the remaining inputs (flowz.hpp:984), but the delay lines of
        output_arity(left) results plus the remaining inputs' lines (:985) -- they differ when the left side passes
        surplus inputs through (:996-999)."""
        o = e.op
        if o == "ph":
            return [ins[e.k - 1]]
        if o == "delay":
            if e.k > len(dins) or dins[e.k - 1] is None:
                raise ValueError("no delay line for this wire")
            return [self._delay(dins[e.k - 1], e.n)]
        if o == "const":
            return [Wire("const", dtype=e.dtype, value=e.value)]
        if o == "param":
            return [Wire("param", k=e.k)]
        if o == "neg":
            return [Wire("neg", a=self._scalar(e.ch[0], ins, dins))]
        if o in ("add", "sub", "mul", "div"):
            return [Wire("op", op=o, a=self._scalar(e.ch[0], ins, dins), b=self._scalar(e.ch[1], ins, dins))]
        l, r = e.ch[0], (e.ch[1] if len(e.ch) > 1 else None)
        if o == "chan":
            return self._walk(l, ins, dins) + self._walk(r, ins, dins)
        if o == "par":
            n = fo.input_arity(l)
            return self._walk(l, ins[:n], dins[:n]) + self._walk(r, ins[n:], dins[n:])
        if o == "seq":
            n_l, n_r = fo.input_arity(l), fo.input_arity(r)
            lo = self._walk(l, ins[:n_l], dins[:n_l])
            node = (lo + [None] * fo.output_arity(l))[:fo.output_arity(l)]
            ro = self._walk(r, lo + ins[n_l:], node + dins[n_l:])
            return ro + lo[n_r:] + ins[n_l + len(lo):]
        if o == "fb":
            cells = [Wire("cell", target=None) for _ in range(fo.output_arity(l))]
            outs = self._walk(l, cells + ins, cells + dins)
            for c, w in zip(cells, outs):
                c.target = w
            return outs
        if o == "bfb":
            n_l, out_l, out_r = fo.input_arity(l), fo.output_arity(l), fo.output_arity(r)
            cells = [Wire("cell", target=None) for _ in range(out_l)]
            res = self._walk(r, cells + ins, cells + dins)
            k = max(n_l - out_r, 0)
            fed = self._walk(l, res + ins[:k], [None] * out_l + dins[:k])
            for c, w in zip(cells, fed):
                c.target = w
            return res
        raise ValueError(o)

    # ---- one tick, by need --------------------------------------------------------------------------------------
    def _value(self, w):
        if w in self.memo:
            return self.memo[w]
        if w.busy:
            raise Loop("feedback loop without a delay")
        w.busy = True
        try:
            k = w.kind
            if k == "in":
                v = self.cur[w.k]
            elif k == "const":
                v = (w.dtype, _NP[w.dtype](w.value))
            elif k == "param":
                v = (fo.F32, np.float32(self.params[w.k]))
            elif k == "cell":
                if w.target is None:
                    raise Loop("a fed-back wire nothing feeds")
                v = self._value(w.target)
            elif k == "delay":
                h = w.src.hist
                v = (fo.F32, h[-w.n] if w.n <= len(h) else np.float32(0))
            elif k == "neg":
                d, x = self._value(w.a)
                v = (d, _NP[d](-x))
            else:
                (da, x), (db, y) = self._value(w.a), self._value(w.b)
                d = max(da, db)
                x, y = _NP[d](x), _NP[d](y)
                with np.errstate(all="ignore"):
                    if w.op == "add": z = x + y
                    elif w.op == "sub": z = x - y
                    elif w.op == "mul": z = x * y
                    elif d == fo.I32: z = 0 if y == 0 else int(x / y)
                    else: z = x / y
                v = (d, _NP[d](z))
        finally:
            w.busy = False
        self.memo[w] = v
        return v

    def tick(self, *xs, dtype=fo.F32):
        self.memo = {}
        dts = [dtype] * len(xs) if isinstance(dtype, int) else list(dtype)
        self.cur = [(d, _NP[d](x)) for x, d in zip(xs, dts)]
        outs = [self._value(w) for w in self.outputs]
        pushed = [np.float32(self._value(w)[1]) for w in self.delay_sources]   # every line is read before it is pushed
        for w, v in zip(self.delay_sources, pushed):
            w.hist.append(v)
        return outs
