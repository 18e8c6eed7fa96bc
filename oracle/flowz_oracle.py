"""CPU oracle for the flowz per-sample evaluator -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under zignal_b200/ does.

What it is: an independent restatement, in Python, of the tick semantics of the reference
(andre-bergner/zignal @ bdc4e29, /root/reference).  It deliberately keeps the reference's *shape* --
a recursive walk over the canonical expression tree with a nested state tree, tuple-style wire
routing, one evaluator function per combinator -- whereas the product lowers the graph to a flat
SSA program.  Agreement between the two is therefore a check of the product's lowering, not a
tautology.  Two back ends share the walk:

  * NumPy: every wire value is an array over channels, ticks are a Python loop (small cases);
  * C emitter: the same walk run once on symbolic values emits a straight-line C tick, compiled
    with `gcc -O3 -ffp-contract=off -fopenmp` (the reference builds with -O3 and neither -march nor
    fast-math, CMakeLists.txt:17-19, so mul and add round separately).  Used for large parity
    cases and as the timed "port" CPU baseline.

Parity pinning: the reference's own tests hold exact known-answer vectors only for small integer
graphs (test/tests.cpp:110-178) and structural expectations for the canonical split
(test/tests.cpp:27-77).  tests/test_oracle_golden.py checks this oracle against all of them.  For
the biquad graphs the reference's tests hold no numbers; the oracle is instead checked bit-for-bit
against the reference's own hand-written biquad lambdas (test/benchmark.cpp:35-126), compiled from
the reference sources where they lie into oracle/_ref (oracle/Makefile).

Where the reference's behaviour is undefined (a delayed read its state sizing does not cover, :950-958 with :1043-1047) the
oracle raises instead of reading a neighbouring slot; graphs the reference cannot compile (unsplittable feedbacks) raise too:
for those the product is checked against tests/netlist_flowz.py.

Reference map (file:line are /root/reference/flowz/flowz.hpp unless noted):
  grammar                        :68-102
  input_arity / output_arity     :162-246
  max/min_input_delays           :286-506
  make_front / add_front_panel   :261-277
  build_state                    :685-725 , to_array :1142-1170
  eval_it dispatch               :740-774
  make_canonical, u2b, split     :794-935
  place_the_holder, place_delay  :941-958
  sequence                       :960-1001
  binary_feedback                :1031-1074
  parallel                       :1076-1101
  rotate_push_back               :130-148
  compile / stateful_lambda      :1181-1249
  ResultType                     :515-644  (result_type(), pinned by test/tests.cpp:182-232)
  tuple_take / tuple_drop        flowz/tuple_tools.hpp:78-160
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------------------
# expression tree
# --------------------------------------------------------------------------------------------------

I32, F32, F64 = 0, 1, 2
C64, C128 = 4, 5                 # std::complex<float> / <double>: result_type() only, never evaluated
_NP = {I32: np.int32, F32: np.float32, F64: np.float64}
_CT = {I32: "int", F32: "float", F64: "double"}


@dataclass(frozen=True)
class Node:
    op: str                      # ph delay const param neg add sub mul div seq par chan fb bfb
    k: int = 0
    n: int = 0
    dtype: int = F32
    value: float = 0.0
    ch: Tuple["Node", ...] = ()

    def __str__(self) -> str:
        o = self.op
        if o == "ph":
            return f"_{self.k}"
        if o == "delay":
            return f"_{self.k}[_{self.n}]"
        if o == "const":
            if self.dtype == I32:
                s = str(int(self.value))
            else:
                s = float(self.value).hex() + ("f" if self.dtype == F32 else "")
            return f"({s})" if s.startswith("-") else s
        if o == "param":
            return f"${self.k}"
        if o == "neg":
            return f"(-{self.ch[0]})"
        if o == "fb":
            return f"(~{self.ch[0]})"
        if o == "bfb":
            return f"bfb({self.ch[0]} , {self.ch[1]})"
        sym = {"add": " + ", "sub": " - ", "mul": "*", "div": "/", "seq": " |= ", "par": " | ", "chan": " , "}[o]
        return f"({self.ch[0]}{sym}{self.ch[1]})"


def ph(k): return Node("ph", k=k)
def seq(a, b): return Node("seq", ch=(a, b))
def par(a, b): return Node("par", ch=(a, b))
def bfb(a, b): return Node("bfb", ch=(a, b))


_TOKEN = re.compile(r"""\s*(?:
      (?P<num>0[xX][0-9a-fA-F]*\.?[0-9a-fA-F]*(?:[pP][+-]?\d+)?f?
            |(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?f?)
    | (?P<ph>_\d+)
    | (?P<par>\$\d+)
    | (?P<cplx>cplxd?\{[^}]*\})
    | (?P<name>bfb|front)
    | (?P<op>\|=|>>|[-+*/~|,()\[\]])
    )""", re.X)


def _tokens(text: str):
    pos, out = 0, []
    text = text.rstrip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f"oracle parse error at {pos}: {text!r}")
        pos = m.end()
        out.append((m.lastgroup, m.group(m.lastgroup)))
    out.append(("end", ""))
    return out


def _literal(tok: str, negate=False) -> Node:
    is_f32 = tok.endswith("f") and not (tok.lower().startswith("0x") and "p" not in tok.lower())
    body = tok[:-1] if is_f32 else tok
    hexa = body.lower().startswith("0x")
    floating = is_f32 or ("." in body) or (hexa and "p" in body.lower()) or (not hexa and ("e" in body.lower()))
    if not floating:
        v = int(body, 0)
        return Node("const", dtype=I32, value=float(-v if negate else v))
    v = float.fromhex(body) if hexa else float(body)
    if negate:
        v = -v
    if is_f32:
        return Node("const", dtype=F32, value=float(np.float32(v)))
    return Node("const", dtype=F64, value=v)


class _Parser:
    """C++ operator precedence: [] > unary > * / > + - > >> > | > |= (right assoc) > ,"""

    def __init__(self, text):
        self.t = _tokens(text)
        self.i = 0

    def peek(self): return self.t[self.i]
    def next(self):
        tok = self.t[self.i]; self.i += 1; return tok
    def accept(self, val):
        if self.t[self.i] == ("op", val):
            self.i += 1; return True
        return False
    def expect(self, val):
        if not self.accept(val):
            raise ValueError(f"oracle parse error: expected {val!r}, got {self.peek()}")

    def comma(self):
        l = self.assign()
        while self.accept(","):
            l = Node("chan", ch=(l, self.assign()))
        return l
    def assign(self):
        l = self.bitor()
        if self.accept("|="):
            return seq(l, self.assign())
        return l
    def bitor(self):
        l = self.shift()
        while self.accept("|"):
            l = par(l, self.shift())
        return l
    def shift(self):
        l = self.additive()
        while self.accept(">>"):
            l = seq(l, self.additive())
        return l
    def additive(self):
        l = self.mult()
        while True:
            if self.accept("+"): l = Node("add", ch=(l, self.mult()))
            elif self.accept("-"): l = Node("sub", ch=(l, self.mult()))
            else: return l
    def mult(self):
        l = self.unary()
        while True:
            if self.accept("*"): l = Node("mul", ch=(l, self.unary()))
            elif self.accept("/"): l = Node("div", ch=(l, self.unary()))
            else: return l
    def unary(self):
        if self.accept("~"): return Node("fb", ch=(self.unary(),))
        if self.accept("+"): return self.unary()
        if self.accept("-"):
            if self.peek()[0] == "num":          # -0.3f is a literal, as for a C++ compiler
                return _literal(self.next()[1], negate=True)
            return Node("neg", ch=(self.unary(),))
        return self.postfix()
    def postfix(self):
        e = self.primary()
        while self.accept("["):
            if e.op != "ph":
                raise ValueError("only placeholders can be delayed")
            kind, val = self.next()
            if (kind, val) == ("op", "-"):
                kind, val = self.next()
            n = int(val[1:]) if kind == "ph" else int(val)
            self.expect("]")
            e = Node("delay", k=e.k, n=n)
        return e
    def primary(self):
        kind, val = self.next()
        if (kind, val) == ("op", "("):
            e = self.comma(); self.expect(")"); return e
        if kind == "ph": return ph(int(val[1:]))
        if kind == "par": return Node("param", k=int(val[1:]))
        if kind == "num": return _literal(val)
        if kind == "cplx":                          # cplx{re,im}: typed, not evaluated
            return Node("const", dtype=C128 if val.startswith("cplxd") else C64, value=0.0)
        if kind == "name" and val == "bfb":
            self.expect("("); l = self.assign(); self.expect(","); r = self.assign(); self.expect(")")
            return bfb(l, r)
        if kind == "name" and val == "front":
            self.expect("("); n = int(self.next()[1]); self.expect(")")
            return make_front(n)
        raise ValueError(f"oracle parse error: unexpected {val!r}")


def parse(text: str) -> Node:
    p = _Parser(text)
    e = p.comma()
    if p.peek()[0] != "end":
        raise ValueError(f"oracle parse error: trailing {p.peek()}")
    return e


# --------------------------------------------------------------------------------------------------
# static analysis (:162-246, :286-506)
# --------------------------------------------------------------------------------------------------

def input_arity(e: Node) -> int:
    o = e.op
    if o in ("ph", "delay"): return e.k
    if o in ("const", "param"): return 0
    if o == "fb": return max(0, input_arity(e.ch[0]) - output_arity(e.ch[0]))
    if o == "bfb":
        l, r = e.ch
        return max(0, input_arity(l) - output_arity(r)) + max(0, input_arity(r) - output_arity(l))
    if o == "par": return input_arity(e.ch[0]) + input_arity(e.ch[1])
    if o == "seq":
        l, r = e.ch
        return input_arity(l) + max(0, input_arity(r) - output_arity(l))
    return max([input_arity(c) for c in e.ch] + [0])


def output_arity(e: Node) -> int:
    o = e.op
    if o in ("chan", "par"): return output_arity(e.ch[0]) + output_arity(e.ch[1])
    if o == "fb": return output_arity(e.ch[0])
    if o == "bfb": return output_arity(e.ch[1])
    if o == "seq":
        l, r = e.ch
        return output_arity(r) + max(0, output_arity(l) - input_arity(r))
    return 1


def _take(t, n): return list(t) if n >= len(t) else list(t[:max(n, 0)])      # tuple_tools.hpp:92-103
def _drop(t, n): return [] if n >= len(t) else list(t[max(n, 0):])           # tuple_tools.hpp:138-150


def _zip_wires(a, b, f):
    m = min(len(a), len(b))
    return [f(x, y) for x, y in zip(a[:m], b[:m])] + list(a[m:]) + list(b[m:])


def _map_min(n, m): return m if n == -1 else n if m == -1 else min(n, m)     # :384-388


def input_delays(e: Node, minimum: bool) -> List[int]:
    other = -1 if minimum else 0
    o = e.op
    if o == "delay": return [other] * (e.k - 1) + [e.n]
    if o == "ph": return [other] * (e.k - 1) + [0]
    if o in ("const", "param"): return []
    if o == "fb": return _drop(input_delays(e.ch[0], minimum), output_arity(e.ch[0]))
    if o == "bfb":
        l, r = e.ch
        return _drop(input_delays(l, minimum), output_arity(r)) + _drop(input_delays(r, minimum), output_arity(l))
    if o == "par": return input_delays(e.ch[0], minimum) + input_delays(e.ch[1], minimum)
    if o == "seq":
        l, r = e.ch
        return input_delays(l, minimum) + _drop(input_delays(r, minimum), output_arity(l))
    acc: List[int] = []
    for c in e.ch:
        acc = _zip_wires(input_delays(c, minimum), acc, _map_min if minimum else max)
    return acc


def max_input_delays(e): return input_delays(e, False)
def min_input_delays(e): return input_delays(e, True)


# --------------------------------------------------------------------------------------------------
# ResultType (:515-644), pinned by test/tests.cpp:182-232
#
# Kept in the reference's two-pass shape: pass 1 walks the flowz expression and, wherever a fed-back wire is
# involved, builds a *type expression* with ABSORBER leaves instead of a type; pass 2 (_resolve) applies the
# absorb_left / absorb_right rules (:536-548) and C++'s usual arithmetic conversions to what pass 1 left over.
# --------------------------------------------------------------------------------------------------

ABSORBER = -1


def _cpp_binary(a: int, b: int) -> int:
    """decltype(a op b) for + - * / on {int, float, double, complex<float>, complex<double>}."""
    cplx = {C64: F32, C128: F64}
    if a in cplx or b in cplx:
        if a in cplx and b in cplx:
            ok = a == b
        elif a in cplx:
            ok = b == cplx[a]
        else:
            ok = a == cplx[b]
        if not ok:
            raise TypeError("no operator for these operand types")        # a compile error in the reference
        return a if a in cplx else b
    return max(a, b)


def _resolve(t):
    """Pass 2: t is a type (int), or ('u', t) / ('b', l, r) left over by pass 1."""
    if isinstance(t, int):
        return t
    if t[0] == "u":
        return _resolve(t[1])                       # "unary-op absorber -> absorber" (:533)
    l, r = t[1], t[2]
    if l == ABSORBER:                               # absorb_left  :539-542  (structural: the operand IS the absorber)
        return _resolve(r)
    if r == ABSORBER:                               # absorb_right :544-545
        return _resolve(l)
    lt, rt_ = _resolve(l), _resolve(r)
    if lt == ABSORBER: return rt_                   # an operand that *resolved* to the absorber: the reference would
    if rt_ == ABSORBER: return lt                   # need another pass; same answer wherever it has one
    return _cpp_binary(lt, rt_)


def _rt1(e: "Node", state: list):
    """Pass 1.  Returns (list of type expressions, is_tuple)."""
    o = e.op
    if o in ("ph", "delay"):                        # get_fn :550-557, :582-589
        if e.k > len(state):
            raise TypeError(f"_{e.k} reads past the typed wires")
        return [state[e.k - 1]], False
    if o == "const": return [e.dtype], False        # :590-593
    if o == "param": return [F32], False
    if o in ("fb", "bfb"):                          # :594-615
        body = e.ch[0] if o == "fb" else e.ch[1]
        n_abs = output_arity(e.ch[0])
        types, _ = _rt1(body, [ABSORBER] * n_abs + list(state))
        return [_resolve(t) for t in types], True   # the second ResultType(...) call + make_flat_tuple
    if o == "seq":                                  # :616-624
        l, _ = _rt1(e.ch[0], state)
        n = input_arity(e.ch[1])
        r, _ = _rt1(e.ch[1], _take(l, n))
        return r + _drop(l, n), True
    if o == "par":                                  # :625-630
        n = input_arity(e.ch[0])
        return _rt1(e.ch[0], _take(state, n))[0] + _rt1(e.ch[1], _drop(state, n))[0], True
    if o == "chan":                                 # :631-636
        return _rt1(e.ch[0], state)[0] + _rt1(e.ch[1], state)[0], True
    kids = []
    for c in e.ch:                                  # proto::_default :637-641
        t, tup = _rt1(c, state)
        if tup:
            raise TypeError("arithmetic on a tuple")
        kids.append(t[0])
    return [("u", kids[0]) if len(kids) == 1 else ("b", kids[0], kids[1])], False


def result_type(expr, in_types: Sequence[int]):
    """(types of the output wires, is_tuple) -- ResultType{}(expr, tuple<in_types...>)."""
    e = parse(expr) if isinstance(expr, str) else expr
    types, tup = _rt1(e, list(in_types))
    return [_resolve(t) for t in types], tup


# --------------------------------------------------------------------------------------------------
# canonical form (:261-277, :794-935)
# --------------------------------------------------------------------------------------------------

def make_front(n: int) -> Node:
    if n < 1:
        raise ValueError("make_front<0> does not exist")
    f = ph(1)
    for _ in range(n - 1):
        f = par(f, ph(1))
    return f


def _is_terminal(e): return e.op in ("ph", "const", "param")


def _needs_num_direct_input(r: Node, n: int) -> bool:          # :670-674, :869-874
    return any(d == 0 for d in _take(min_input_delays(r), n))


def _u2b(e: Node) -> List[Node]:                                # :811-846
    if _is_terminal(e):
        return [e]
    if e.op == "seq":
        return _split_in_sequence(e.ch[0], e.ch[1])
    res = [_u2b(c) for c in e.ch]
    rebuilt = Node(e.op, e.k, e.n, e.dtype, e.value, tuple(r[0] for r in res))
    return [rebuilt] + (res[0][1:] if res else [])


def _split_in_sequence(l: Node, r: Node) -> List[Node]:         # :887-935
    ul = _u2b(l)
    if len(ul) == 2:
        return [ul[0], seq(ul[1], r)]
    if _needs_num_direct_input(r, output_arity(ul[0])):
        ur = _u2b(r)
        if len(ur) == 2:
            return [seq(ul[0], ur[0]), ur[1]]
        return [seq(ul[0], ur[0])]
    return [ul[0], r]


def make_canonical(e: Node) -> Node:                            # :794-805, :862-884
    if _is_terminal(e):
        return e
    if e.op == "fb":
        x = e.ch[0]
        parts = _u2b(make_canonical(seq(make_front(output_arity(x)), x)))
        if len(parts) != 2:
            raise ValueError(f"feedback cannot be split: {e}")
        return bfb(parts[0], parts[1])
    return Node(e.op, e.k, e.n, e.dtype, e.value, tuple(make_canonical(c) for c in e.ch))


def canonical_with_front(e: Node) -> Node:                      # compile() :1233-1249
    n = input_arity(e)
    if n == 0:                                                  # extension: no make_front<0> in the reference
        return make_canonical(e)
    return make_canonical(seq(make_front(n), e))


# --------------------------------------------------------------------------------------------------
# state tree (:685-725).  A delay line is the pair [depth, slot]; `slot` is whatever the back end
# uses to find the storage.  None stands for no_state (:128, :1148-1152).
# --------------------------------------------------------------------------------------------------

class _Alloc:
    def __init__(self): self.n = 0
    def line(self, depth):
        if depth == 0:
            return None
        off = self.n
        self.n += depth
        return (depth, off)


def build_state(e: Node, alloc: _Alloc):
    if e.op in ("seq", "bfb"):
        l, r = e.ch
        node_state = [alloc.line(d) for d in _take(max_input_delays(r), output_arity(l))]
        return [node_state, build_state(l, alloc), build_state(r, alloc)]
    if e.op == "par":
        return [build_state(e.ch[0], alloc), build_state(e.ch[1], alloc)]
    if e.op == "fb":
        raise ValueError("unary feedback must be canonicalised first")
    return []


# --------------------------------------------------------------------------------------------------
# the tick, generic over a back end
# --------------------------------------------------------------------------------------------------

class _Bottom:                                                  # bottom_type :1004
    pass


BOTTOM = _Bottom()


def _flatten(x):                                                # flatten_tuple, tuple_tools.hpp:65-70
    if isinstance(x, list):
        out = []
        for y in x:
            out.extend(_flatten(y))
        return out
    return [x]


class _Walker:
    """eval_it (:740-774) and the evaluators it dispatches to."""

    def __init__(self, backend):
        self.b = backend

    def eval(self, e: Node, inp: list, delayed):
        in_state, my_state = delayed
        o = e.op
        if o == "delay":                                        # place_delay :950-958
            line = in_state[e.k - 1]
            if line is None:
                raise ValueError(f"{e}: wire has no delay line")
            return self.b.read(line, e.n)
        if o == "ph":                                           # place_the_holder :941-948
            return inp[e.k - 1]
        if o == "const": return self.b.const(e.dtype, e.value)
        if o == "param": return self.b.param(e.k)
        if o == "bfb": return self.binary_feedback(e.ch[0], e.ch[1], inp, delayed)
        if o == "seq": return self.sequence(e.ch[0], e.ch[1], inp, delayed)
        if o == "par": return self.parallel(e.ch[0], e.ch[1], inp, delayed)
        if o == "chan":                                         # :765-768, same environment for both
            return [self.eval(e.ch[0], inp, delayed), self.eval(e.ch[1], inp, delayed)]
        if o == "neg":
            return self.b.neg(self._scalar(e.ch[0], inp, delayed))
        if o in ("add", "sub", "mul", "div"):                   # proto::_default :769-772
            a = self._scalar(e.ch[0], inp, delayed)
            c = self._scalar(e.ch[1], inp, delayed)
            return self.b.arith(o, a, c)
        raise ValueError(f"cannot evaluate {o}")

    def _scalar(self, e, inp, delayed):
        v = self.eval(e, inp, delayed)
        if isinstance(v, list) or v is BOTTOM:
            raise ValueError(f"arithmetic operand is not a single value: {e}")
        return v

    def sequence(self, l, r, inp, delayed):                     # :960-1001
        in_state, (node_state, left_state, right_state) = delayed
        n_l = input_arity(l)
        left_result = _flatten([self.eval(l, _take(inp, n_l), (_take(in_state, n_l), left_state))])
        right_input = left_result + _drop(inp, n_l)
        right_delayed = node_state + _drop(in_state, n_l)
        right_result = _flatten([self.eval(r, right_input, (right_delayed, right_state))])
        for line, y in zip(node_state, left_result):            # tuple_for_each over the shorter
            self.b.push(line, y)
        return right_result + _drop(left_result, input_arity(r)) + _drop(inp, n_l + len(left_result))

    def binary_feedback(self, l, r, inp, delayed):              # :1031-1074
        in_state, (node_state, left_state, right_state) = delayed
        n_extra = input_arity(l) - output_arity(r)
        if n_extra < 0:
            raise ValueError("binary_feedback: ill-formed (promise part too narrow)")
        future_input = [BOTTOM] * output_arity(l) + list(inp)   # tuple_drop<min(0, ..)> == drop<0>
        result = _flatten([self.eval(r, future_input, (node_state + list(in_state), right_state))])
        promise_input = result + _take(inp, n_extra)
        promise_delayed = [None] * output_arity(l) + _take(in_state, n_extra)
        promise_result = _flatten([self.eval(l, promise_input, (promise_delayed, left_state))])
        for line, y in zip(node_state, promise_result):
            self.b.push(line, y)
        return result

    def parallel(self, l, r, inp, delayed):                     # :1076-1101
        in_state, (left_state, right_state) = delayed
        n_l = input_arity(l)
        return [self.eval(l, _take(inp, n_l), (_take(in_state, n_l), left_state)),
                self.eval(r, _drop(inp, n_l), (_drop(in_state, n_l), right_state))]


def _promote(a, b): return max(a, b)


# binary_feedback hands its future part ALL external inputs (:1043-1047, the drop<min(0, ...)> with its TODO), while
# the delay analysis that sizes the state (:443-506) gives the future part the wires after the promise part's.  Where the
# two disagree a delayed read indexes before the start of its std::array (s[size - n] with n > size, :950-958):
# undefined behaviour in the reference, an error here and in the product.
_OUT_OF_LINE = "delayed read reaches past the delay line the reference allocates (out-of-bounds in the reference)"


class _NumpyBackend:
    """Values are (dtype, ndarray[C]).  Pushes are deferred to the end of the tick, which is
    equivalent because every line is read before its owner pushes it."""

    def __init__(self, n_state, channels, params):
        self.state = np.zeros((n_state, channels), np.float32)
        self.params = params
        self.channels = channels
        self.pending = []

    def read(self, line, n):                                    # s[s.size() - n]
        depth, off = line
        if n > depth:
            raise ValueError(_OUT_OF_LINE)
        return (F32, self.state[off + depth - n].copy())
    def const(self, dtype, value):
        return (dtype, np.full(self.channels, value, _NP[dtype]))
    def param(self, k):
        return (F32, np.broadcast_to(np.asarray(self.params[k], np.float32), (self.channels,)).copy())
    def neg(self, a):
        return (a[0], -a[1])
    def arith(self, op, a, c):
        dt = _promote(a[0], c[0])
        x, y = a[1].astype(_NP[dt]), c[1].astype(_NP[dt])
        with np.errstate(all="ignore"):
            if op == "add": z = x + y
            elif op == "sub": z = x - y
            elif op == "mul": z = x * y
            elif dt == I32: z = np.where(y == 0, 0, np.trunc(x / np.where(y == 0, 1, y))).astype(np.int32)
            else: z = x / y
        return (dt, z.astype(_NP[dt]))
    def push(self, line, y):
        if line is None:
            return
        if y is BOTTOM:
            raise ValueError("pushing an unresolved fed-back wire")
        self.pending.append((line, y[1].astype(np.float32)))    # narrowing to the float line :136
    def end_tick(self):
        for (depth, off), y in self.pending:                    # rotate_push_back :130-148
            self.state[off:off + depth - 1] = self.state[off + 1:off + depth]
            self.state[off + depth - 1] = y
        self.pending = []


class Oracle:
    """compile() + stateful_lambda for `channels` independent voices (NumPy back end)."""

    def __init__(self, expr: str, channels: int = 1, params: Optional[Sequence] = None):
        self.user = parse(expr)
        self.n_in = input_arity(self.user)
        self.n_out = output_arity(self.user)
        self.canonical = canonical_with_front(self.user)
        alloc = _Alloc()
        self.state_tree = build_state(self.canonical, alloc)
        self.n_state = alloc.n
        self.channels = channels
        self.backend = _NumpyBackend(self.n_state, channels, list(params or []))
        self.walker = _Walker(self.backend)

    def tick(self, *xs, dtype=F32):
        """xs: n_in scalars or arrays[C]; dtype: the C++ type of the arguments (one for all, or one per argument).
        Returns list of (dtype, array[C])."""
        if len(xs) != self.n_in:
            raise ValueError("wrong number of inputs")
        dts = [dtype] * len(xs) if isinstance(dtype, int) else list(dtype)
        inp = [(d, np.broadcast_to(np.asarray(x, _NP[d]), (self.channels,)).copy()) for x, d in zip(xs, dts)]
        res = _flatten([self.walker.eval(self.canonical, inp, ([], self.state_tree))])   # :1193-1201
        self.backend.end_tick()
        return res

    def process(self, inputs: Sequence[np.ndarray]) -> List[np.ndarray]:
        """inputs: n_in arrays [C, T] float32 -> n_out arrays [C, T] float32."""
        T = inputs[0].shape[1] if self.n_in else 0
        outs = [np.zeros((self.channels, T), np.float32) for _ in range(self.n_out)]
        for t in range(T):
            res = self.tick(*[x[:, t] for x in inputs])
            for o, (_, y) in zip(outs, res):
                o[:, t] = y.astype(np.float32)
        return outs


# --------------------------------------------------------------------------------------------------
# C emitter: same walk, symbolic values
# --------------------------------------------------------------------------------------------------

class _CBackend:
    def __init__(self):
        self.lines: List[str] = []
        self.n = 0
        self.pending = []

    def _new(self, dtype, rhs):
        name = f"v{self.n}"
        self.n += 1
        self.lines.append(f"const {_CT[dtype]} {name} = {rhs};")
        return (dtype, name)
    def read(self, line, n):
        depth, off = line
        if n > depth:
            raise ValueError(_OUT_OF_LINE)
        return self._new(F32, f"s[{off + depth - n}]")
    def const(self, dtype, value):
        if dtype == I32: return self._new(I32, str(int(value)))
        if dtype == F32: return self._new(F32, float(value).hex() + "f")
        return self._new(F64, float(value).hex())
    def param(self, k): return self._new(F32, f"p[{k}]")
    def neg(self, a): return self._new(a[0], f"-{a[1]}")
    def arith(self, op, a, c):
        dt = _promote(a[0], c[0])
        sym = {"add": "+", "sub": "-", "mul": "*", "div": "/"}[op]
        return self._new(dt, f"({_CT[dt]}){a[1]} {sym} ({_CT[dt]}){c[1]}")
    def push(self, line, y):
        if line is None:
            return
        if y is BOTTOM:
            raise ValueError("pushing an unresolved fed-back wire")
        self.pending.append((line, y))
    def end_tick(self):
        for (depth, off), y in self.pending:
            for j in range(depth - 1):
                self.lines.append(f"s[{off + j}] = s[{off + j + 1}];")
            self.lines.append(f"s[{off + depth - 1}] = (float){y[1]};")
        self.pending = []


def emit_c(expr: str) -> Tuple[str, int, int, int]:
    """Straight-line C for one tick of `expr`, wrapped in a channel/time loop.
    Returns (source, n_in, n_out, n_state).  Layout: planar [C][ld]; state [C][n_state]."""
    user = parse(expr)
    n_in, n_out = input_arity(user), output_arity(user)
    canonical = canonical_with_front(user)
    alloc = _Alloc()
    tree = build_state(canonical, alloc)
    b = _CBackend()
    inp = [(F32, f"x{i}") for i in range(n_in)]
    res = _flatten([_Walker(b).eval(canonical, inp, ([], tree))])
    if any(v is BOTTOM for v in res):
        raise ValueError("an output is a fed-back wire nothing ever assigns (bottom_type, :1004)")
    n_out = len(res)        # the tick decides: `sequence` passes surplus inputs through (:996-999), output_arity
                            # (:238-247) does not count them -- `_1 |= (_1[_3] | _2[_1])` returns three values
    b.end_tick()
    body = "\n            ".join(b.lines)
    loads = "\n            ".join(f"const float x{i} = in[{i}][c * ld_in + t];" for i in range(n_in))
    stores = "\n            ".join(f"out[{j}][c * ld_out + t] = (float){v[1]};" for j, v in enumerate(res))
    src = f"""// generated by oracle/flowz_oracle.py emit_c() -- oracle only, never shipped
// expr: {expr}
#include <stddef.h>
void zg_oracle_run(const float* const* in, float* const* out, long channels, long n_samples,
                   long ld_in, long ld_out, float* state, const float* params, long param_stride)
{{
    #pragma omp parallel for schedule(static)
    for (long c = 0; c < channels; ++c) {{
        float s[{max(alloc.n, 1)}];
        const float* p = params + c * param_stride;
        for (int i = 0; i < {alloc.n}; ++i) s[i] = state[c * {alloc.n} + i];
        for (long t = 0; t < n_samples; ++t) {{
            {loads}
            {body}
            {stores}
        }}
        for (int i = 0; i < {alloc.n}; ++i) state[c * {alloc.n} + i] = s[i];
    }}
}}
"""
    return src, n_in, n_out, alloc.n


_HERE = os.path.dirname(os.path.abspath(__file__))


class COracle:
    """The emitted C tick, compiled and loaded.  process() matches Oracle.process() bit for bit."""

    def __init__(self, expr: str, channels: int, params: Optional[np.ndarray] = None, threads: Optional[int] = None):
        src, self.n_in, self.n_out, self.n_state = emit_c(expr)
        build = os.path.join(_HERE, "_build")
        os.makedirs(build, exist_ok=True)
        tag = hashlib.sha1(src.encode()).hexdigest()[:16]
        so = os.path.join(build, f"oracle_{tag}.so")
        if not os.path.exists(so):
            cfile = os.path.join(build, f"oracle_{tag}.c")
            with open(cfile, "w") as f:
                f.write(src)
            tmp = so + f".tmp{os.getpid()}"
            subprocess.check_call(["gcc", "-O3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
                                   "-o", tmp, cfile])
            os.replace(tmp, so)
        self.lib = ctypes.CDLL(so)
        self.channels = channels
        self.state = np.zeros((channels, max(self.n_state, 1)), np.float32)
        if params is None:
            params = np.zeros((channels, 1), np.float32)
        params = np.ascontiguousarray(params, np.float32)
        if params.ndim == 1:                                    # scalar per parameter, broadcast
            self.params, self.pstride = params, 0
        else:                                                   # [C, n_params]
            self.params, self.pstride = params, params.shape[1]
        self.threads = threads

    def process(self, inputs: Sequence[np.ndarray]) -> List[np.ndarray]:
        C = self.channels
        T = inputs[0].shape[1] if self.n_in else 0
        ins = [np.ascontiguousarray(x, np.float32) for x in inputs]
        outs = [np.zeros((C, T), np.float32) for _ in range(self.n_out)]
        P = ctypes.POINTER(ctypes.c_float)
        in_arr = (P * max(self.n_in, 1))(*[x.ctypes.data_as(P) for x in ins])
        out_arr = (P * max(self.n_out, 1))(*[y.ctypes.data_as(P) for y in outs])
        if self.threads:
            os.environ["OMP_NUM_THREADS"] = str(self.threads)
        self.lib.zg_oracle_run(in_arr, out_arr, ctypes.c_long(C), ctypes.c_long(T), ctypes.c_long(T),
                               ctypes.c_long(T), self.state.ctypes.data_as(P),
                               self.params.ctypes.data_as(P), ctypes.c_long(self.pstride))
        return outs


# --------------------------------------------------------------------------------------------------
# workload helpers shared by tests and bench (inputs, coefficients); see SURVEY.md section 8(d)
# --------------------------------------------------------------------------------------------------

def noise(channels: int, samples: int, seed: int = 0) -> np.ndarray:
    """x[c,t] ~ U(-1,1) from a counter hash of (seed, c, t): identical wherever it is generated."""
    c = np.arange(channels, dtype=np.uint64)[:, None]
    t = np.arange(samples, dtype=np.uint64)[None, :]
    sd = np.uint64((int(seed) * 0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF)      # wraps, like the uint64 product
    h = (c * np.uint64(0x9E3779B97F4A7C15) + t * np.uint64(0xC2B2AE3D27D4EB4F) + sd) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(33); h *= np.uint64(0xFF51AFD7ED558CCD)
    h ^= h >> np.uint64(33); h *= np.uint64(0xC4CEB9FE1A85EC53)
    h ^= h >> np.uint64(33)
    u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))   # [0,1), 24 bits
    return (u * np.float32(2.0) - np.float32(1.0)).astype(np.float32)


def rbj_lowpass(f: float, q: float = 0.707, sr: float = 44100.0):
    """RBJ low-pass section in flowz sign convention y = b0 x + b1 x1 + b2 x2 + a1 y1 + a2 y2
    (test/benchmark.cpp:25-26); formulae as in reactive_equations/reactive_filter_coeff.cpp:16-50."""
    w0 = 2.0 * np.pi * f / sr
    alpha = np.sin(w0) / (2.0 * q)
    a0 = 1.0 + alpha
    b1 = (1.0 - np.cos(w0)) / a0
    b0 = b1 / 2.0
    return (np.float32(b0), np.float32(b1), np.float32(b0),
            np.float32(2.0 * np.cos(w0) / a0), np.float32(-(1.0 - alpha) / a0))


def lit(x) -> str:
    """exact float literal"""
    return float(np.float32(x)).hex() + "f"


def biquad_df1(b0, b1, b2, a1, a2) -> str:
    """fwd |= bwd of test/benchmark.cpp:25-33"""
    return (f"({lit(b0)}*_1 + {lit(b1)}*_1[_1] + {lit(b2)}*_1[_2]"
            f" |= ~(_2 + {lit(a1)}*_1[_1] + {lit(a2)}*_1[_2]))")


def biquad_cascade(sections: int = 4) -> str:
    """`sections` stable RBJ low-pass DF1 sections in series, f = 440 * 2^k Hz; the
    octaves wrap after six sections so that every cutoff stays below Nyquist (22 050 Hz)."""
    return " |= ".join(biquad_df1(*rbj_lowpass(440.0 * 2 ** (k % 6))) for k in range(sections))


def biquad_cascade_f64(x: np.ndarray, sections: int = 4) -> np.ndarray:
    """The same cascade evaluated in float64 (direct form 1, same fp32 coefficients): the yardstick
    for how far fp32 rounding alone moves the reference's own output (tests of the FMA mode)."""
    y = np.asarray(x, np.float64)
    for k in range(sections):
        b0, b1, b2, a1, a2 = (float(v) for v in rbj_lowpass(440.0 * 2 ** (k % 6)))
        x1 = x2 = y1 = y2 = np.zeros(y.shape[0])
        out = np.empty_like(y)
        for t in range(y.shape[1]):
            xt = y[:, t]
            yt = ((b0 * xt + b1 * x1) + b2 * x2 + a1 * y1) + a2 * y2
            x2, x1, y2, y1 = x1, xt, y1, yt
            out[:, t] = yt
        y = out
    return y


def biquad_cascade_params(sections: int = 4) -> str:
    """Same cascade with every coefficient a run-time parameter $0..$(5*sections-1)."""
    parts = []
    for k in range(sections):
        p = 5 * k
        parts.append(f"(${p}*_1 + ${p+1}*_1[_1] + ${p+2}*_1[_2] |= ~(_2 + ${p+3}*_1[_1] + ${p+4}*_1[_2]))")
    return " |= ".join(parts)


# ---- FIR (BASELINE configs[3]; SURVEY.md 8d: windowed-sinc taps shared by all channels) ----------

def fir_taps(n: int = 256, cutoff: float = 0.25) -> np.ndarray:
    """n-tap Hamming-windowed sinc low-pass (cutoff in cycles/sample * 2), DC gain ~ 1, fp32."""
    k = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    h = np.sinc(k * cutoff) * np.hamming(n) if n > 1 else np.ones(1)
    return (h / h.sum()).astype(np.float32)


def fir_expr(taps: Sequence[float]) -> str:
    """c0*_1 + c1*_1[_1] + ... in the reference's spelling (flowz.hpp:84-85 delays, :769-772 arithmetic);
    C++ associates the sum to the left."""
    return " + ".join(f"{lit(c)}*_1" if k == 0 else f"{lit(c)}*_1[_{k}]" for k, c in enumerate(taps))


def fir_expr_params(n: int) -> str:
    """Same FIR with every tap a run-time parameter $0..$(n-1)."""
    return " + ".join(f"${k}*_1" if k == 0 else f"${k}*_1[_{k}]" for k in range(n))


def fir_direct(x: np.ndarray, taps: Sequence[float], history: Optional[np.ndarray] = None) -> np.ndarray:
    """What the tick of fir_expr(taps) computes, vectorised over the block: every product and every sum
    rounded to fp32 separately, summed in tap order (numpy float32 arithmetic does exactly that).
    `history` = [C][n-1] samples before the block (the delay line, oldest first), zeros if None.
    Checked against the tick-by-tick oracles in tests/test_oracle_golden.py."""
    x = np.asarray(x, np.float32)
    C_, T = x.shape
    n = len(taps)
    h = np.zeros((C_, n - 1), np.float32) if history is None else np.asarray(history, np.float32)
    ext = np.concatenate([h, x], axis=1)
    acc = np.float32(taps[0]) * ext[:, n - 1:n - 1 + T]
    for k in range(1, n):
        acc = acc + np.float32(taps[k]) * ext[:, n - 1 - k:n - 1 - k + T]
    return acc


# ---- bf16 sample storage (BASELINE configs[4]) and the polyphonic voice graph -----------------------

def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16 -> fp32 (finite inputs): what storing a sample as bf16 does."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """the 16 stored bits of bf16_round(x)"""
    return (bf16_round(x).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def osc_expr(f: float = 440.0, sr: float = 44100.0) -> str:
    """Recursive sine oscillator y = k*y1 - y2 + x, dirac-excited (flowz has no sin, TODO.md:4;
    zero-input graphs do not compile, TODO.md:65); k = 2 cos(2 pi f / sr)."""
    return f"~({lit(2.0 * np.cos(2.0 * np.pi * f / sr))}*_1[_1] - _1[_2] + _2)"


def osc_lp_expr(f: float = 440.0, a: float = 0.9) -> str:
    """BASELINE configs[2]: sine oscillator >> one-pole low-pass."""
    return f"{osc_expr(f)} |= ~(_2 + {lit(a)}*_1[_1])"


def poly_voice_expr(f: float = 440.0, g: float = 0.25) -> str:
    """BASELINE configs[4]: osc >> biquad >> (biquad inside a unit-delayed feedback
    loop of gain g):  osc |= (fwd|=bwd) |= ~( (_2 + g*_1[_1]) |= (fwd|=bwd) )."""
    bq1 = biquad_df1(*rbj_lowpass(1760.0))
    bq2 = biquad_df1(*rbj_lowpass(3520.0))
    return f"{osc_expr(f)} |= {bq1} |= ~((_2 + {lit(g)}*_1[_1]) |= {bq2})"
