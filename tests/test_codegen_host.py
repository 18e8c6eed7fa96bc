"""The device code generator checked without a device: the tick functor that zg_codegen.cpp writes for the NVRTC-built
kernels (EXACT mode: __fadd_rn / __fmul_rn / ..., one statement per SSA node, state pushes at the end) is extracted
from the generated CUDA source, compiled for the HOST with those intrinsics defined as the plain IEEE operations
(g++ -ffp-contract=off), and ticked over noise in two blocks.  It must reproduce the oracle bit for bit -- and, for the
graphs the reference cannot compile (kept-whole feedbacks), the netlist evaluator.  What this does not cover is the
streaming skeleton around the functor (TMA, shared memory, state rows): that is what the -m gpu tests are for."""
import ctypes
import os
import random
import re
import subprocess

import numpy as np
import pytest

import flowz_oracle as fo
import netlist_flowz as nl
from test_fuzz_frontend import _gen

HOST_PRELUDE = r"""
#include <cstring>
#define __device__
#define __forceinline__ inline
#define ZG_SYNTH_MASK 0u
namespace zgk {
template <int N> struct Arr {
    float v[N > 0 ? N : 1];
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
}
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
struct Params { const float* p; float operator[](int i) const { return p[i]; } };
// ticks that carry delayed products declare N_EXTRA registers, rebuilt from the state at block start (init)
template <class T, class = void> struct extra_count { static constexpr int value = 0; };
template <class T> struct extra_count<T, decltype((void)T::N_EXTRA)> { static constexpr int value = T::N_EXTRA; };
int g_extra_seen = 0;
template <class Tick>
static void run_tick(const float* const* in, float* const* out, long n, float* state, const float* params) {
    zgk::Arr<Tick::N_IN> x;
    zgk::Arr<Tick::N_OUT> y;
    zgk::Arr<Tick::N_STATE> s;
    for (int j = 0; j < Tick::N_STATE; ++j) s[j] = state[j];
    const Params p{params};
    constexpr int NE = extra_count<Tick>::value;
    zgk::Arr<NE> e;
    if constexpr (NE > 0) { Tick::init(s, p, e); g_extra_seen += 1; }
    for (long t = 0; t < n; ++t) {
        for (int k = 0; k < Tick::N_IN; ++k) x[k] = in[k][t];
        if constexpr (NE > 0) Tick::tick(x, y, s, p, e);
        else Tick::tick(x, y, s, p);
        for (int o = 0; o < Tick::N_OUT; ++o) out[o][t] = y[o];
    }
    for (int j = 0; j < Tick::N_STATE; ++j) state[j] = s[j];
}
extern "C" int zg_host_extra_seen() { return g_extra_seen; }
"""


def _tick_struct(zg, expr, name):
    src = zg.compile(expr).kernel(cubin=False).decode()
    m = re.search(r"struct ZgTick \{.*?\n\};\n", src, re.S)
    assert m, "generated source has no tick functor"
    return m.group(0).replace("struct ZgTick", f"struct {name}")


def _graphs(zg):
    rng = random.Random(4242)
    consts = ["0.5f", "0.25f", "-0.75f", "0x1p-1f", "1.5f", "-1.0f"]
    out = ["~(_2 + 0.9f*_1[_1])", fo.osc_lp_expr(), fo.poly_voice_expr(), fo.biquad_cascade(3),
           fo.biquad_cascade_params(2), "$0*_1 + $1*_1[_2] |= ~(_2 + $2*_1[_1])",
           "0.5f*_1 + 0.25f*_1[_1] + 0.5f*_1[_2] + 0.25f*_1[_3]", "$0*_1 + $1*_1[_1] + $0*_1[_2] |= ~(_2 + $1*_1[_1])",
           "~~( _1 + _2 + 1.0f |= _1[_1] )", "~( (0.5f*_1 + _2) |= ~(_1 + _2 |= _1[_1]) )", "~((_1[_2] |= _2) | (_1 - _2))"]
    while len(out) < 60:
        e = _gen(rng, rng.randint(2, 5), rng.randint(1, 3), consts=consts)
        try:
            g = zg.compile(e)
        except zg.ZgError:
            continue
        if g.all_f32 and g.n_in >= 1 and g.n_out >= 1 and g.n_state <= 64 and ("~" in e or "|" in e or "," in e):
            out.append(e)
    return out


def test_generated_tick_functors_on_the_host(zg, tmp_path):
    exprs = _graphs(zg)
    parts = [HOST_PRELUDE]
    for i, e in enumerate(exprs):
        parts.append(_tick_struct(zg, e, f"Tick{i}"))
    parts.append('extern "C" void zg_host_run(int which, const float* const* in, float* const* out, long n, float* state, '
                 "const float* params) {\n    switch (which) {\n" +
                 "".join(f"        case {i}: run_tick<Tick{i}>(in, out, n, state, params); break;\n" for i in range(len(exprs))) +
                 "    }\n}\n")
    cpp, so = tmp_path / "ticks.cpp", tmp_path / "ticks.so"
    cpp.write_text("\n".join(parts))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(so), str(cpp)])
    lib = ctypes.CDLL(str(so))
    P = ctypes.POINTER(ctypes.c_float)
    T, extended = 96, 0
    for i, e in enumerate(exprs):
        g = zg.compile(e)
        x = [np.ascontiguousarray(fo.noise(1, T, seed=50 + 7 * i + k)[0]) for k in range(g.n_in)]
        y = [np.zeros(T, np.float32) for _ in range(g.n_out)]
        state = np.zeros(max(g.n_state, 1), np.float32)
        params = (np.random.default_rng(i).uniform(-0.6, 0.6, max(g.n_params, 1))).astype(np.float32)   # std::ref terminals
        for t0, n in ((0, 41), (41, T - 41)):                        # two blocks: the state array carries over
            ins = (P * max(g.n_in, 1))(*[a[t0:].ctypes.data_as(P) for a in x])
            outs = (P * g.n_out)(*[a[t0:].ctypes.data_as(P) for a in y])
            lib.zg_host_run(i, ins, outs, ctypes.c_long(n), state.ctypes.data_as(P), params.ctypes.data_as(P))
        try:
            want = [w[0] for w in fo.COracle(e, 1, params=params).process([a[None, :] for a in x])]
        except ValueError:                                           # beyond the reference: the netlist is the checker
            net = nl.Netlist(e, params=params)
            ticks = [net.tick(*[float(a[t]) for a in x]) for t in range(T)]
            want = [np.array([tk[o][1] for tk in ticks], np.float32) for o in range(g.n_out)]
            extended += 1
        assert len(want) == g.n_out, e
        for o in range(g.n_out):
            same = (y[o].view(np.uint32) == want[o].view(np.uint32)) | (np.isnan(y[o]) & np.isnan(want[o]))
            assert same.all(), f"{e}: output {o} differs at {np.argwhere(~same)[:3].ravel().tolist()}"
    assert extended >= 3
    assert lib.zg_host_extra_seen() >= 2 * 4          # the delayed-product reuse was exercised (four graphs, two blocks each)
