// zignal-b200 :: prebuilt tick for cascades of direct-form-1 biquads (K1).
//
// The graph (reference spelling, test/benchmark.cpp:25-33, chained SECTIONS times with |=):
//     fwd = b0*_1 + b1*_1[_1] + b2*_1[_2]
//     bwd = ~( _2 + a1*_1[_1] + a2*_1[_2] )
//     section = fwd |= bwd
// Tick arithmetic per section, in the association the C++ parse tree gives:
//     v = (b0*x + b1*x1) + b2*x2 ;  y = (v + a1*y1) + a2*y2
// State layout = the layout the lowering produces for this graph after line sharing (the y line of
// section k is the x line of section k+1):  s[2k] = signal k two ticks ago, s[2k+1] = one tick ago,
// signal 0 = input, signal k = output of section k.  Parameters: p[5k..5k+4] = b0 b1 b2 a1 a2.
//
// kExact: every product and sum is rounded separately (__fmul_rn/__fadd_rn are never contracted)
//         -> bit-identical to the reference built without FMA (CMakeLists.txt:17-19).
// else:   FMA contraction of the same association (<= 1e-5 block-relative vs the reference).
// kSym (EXACT, coefficients shared by all channels): every section has b0 == b2 bit for bit (all RBJ low-,
//         high-pass and notch sections do).  Then b2*x[t-2] IS the product b0*x computed two ticks earlier
//         -- same operands, same rounding -- so it is carried in a register instead of being recomputed:
//         8 instead of 9 instructions per section per sample, and the EXACT kernel is issue-bound.
#pragma once
#include "zg_stream.cuh"

namespace zgk {

template <int SECTIONS, bool kExact, bool kSym = false>
struct BiquadDf1Cascade {
    static_assert(!kSym || kExact, "product reuse is an EXACT-mode optimisation (FMA folds the product away)");
    static constexpr int N_IN = 1, N_OUT = 1;
    static constexpr int N_STATE = 2 * (SECTIONS + 1);
    static constexpr int N_PARAM = 5 * SECTIONS;
    static constexpr int N_EXTRA = kSym ? 2 * SECTIONS : 0;   // registers that are not delay-line state
    static constexpr unsigned SYNTH_MASK = 0;
#ifdef ZG_K1_CHUNK_UNROLL
    static constexpr int CHUNK_UNROLL = ZG_K1_CHUNK_UNROLL;
#endif

    // e[2k] = b0*x[t-2], e[2k+1] = b0*x[t-1] of section k, rebuilt from the delay lines at block start
    template <class P>
    static __device__ __forceinline__ void init(const Arr<N_STATE>& s, const P& p, Arr<N_EXTRA>& e) {
#pragma unroll
        for (int k = 0; k < (kSym ? SECTIONS : 0); ++k) {
            e[2 * k] = __fmul_rn(p[5 * k], s[2 * k]);
            e[2 * k + 1] = __fmul_rn(p[5 * k], s[2 * k + 1]);
        }
    }
    template <class P>
    static __device__ __forceinline__ void tick(const Arr<N_IN>& x, Arr<N_OUT>& y, Arr<N_STATE>& s,
                                                const P& p, Arr<N_EXTRA>& e) {
        float sig[SECTIONS + 1];
        sig[0] = x[0];
#pragma unroll
        for (int k = 0; k < SECTIONS; ++k) {
            const float b0 = p[5 * k], b1 = p[5 * k + 1], a1 = p[5 * k + 3], a2 = p[5 * k + 4];
            const float x1 = s[2 * k + 1];
            const float y1 = s[2 * k + 3], y2 = s[2 * k + 2];
            const float q0 = __fmul_rn(b0, sig[k]);
            const float v = __fadd_rn(__fadd_rn(q0, __fmul_rn(b1, x1)), e[2 * k]);
            sig[k + 1] = __fadd_rn(__fadd_rn(v, __fmul_rn(a1, y1)), __fmul_rn(a2, y2));
            e[2 * k] = e[2 * k + 1];
            e[2 * k + 1] = q0;
        }
        y[0] = sig[SECTIONS];
#pragma unroll
        for (int k = 0; k <= SECTIONS; ++k) {
            s[2 * k] = s[2 * k + 1];
            s[2 * k + 1] = sig[k];
        }
    }

    template <class P>
    static __device__ __forceinline__ void tick(const Arr<N_IN>& x, Arr<N_OUT>& y, Arr<N_STATE>& s,
                                                const P& p) {
        float sig[SECTIONS + 1];
        sig[0] = x[0];
#pragma unroll
        for (int k = 0; k < SECTIONS; ++k) {
            const float b0 = p[5 * k], b1 = p[5 * k + 1], b2 = p[5 * k + 2], a1 = p[5 * k + 3], a2 = p[5 * k + 4];
            const float x1 = s[2 * k + 1], x2 = s[2 * k];
            const float y1 = s[2 * k + 3], y2 = s[2 * k + 2];
            if (kExact) {
                const float v = __fadd_rn(__fadd_rn(__fmul_rn(b0, sig[k]), __fmul_rn(b1, x1)), __fmul_rn(b2, x2));
                sig[k + 1] = __fadd_rn(__fadd_rn(v, __fmul_rn(a1, y1)), __fmul_rn(a2, y2));
            } else {
                const float v = fmaf(b2, x2, fmaf(b1, x1, b0 * sig[k]));
                sig[k + 1] = fmaf(a2, y2, fmaf(a1, y1, v));
            }
        }
        y[0] = sig[SECTIONS];
#pragma unroll
        for (int k = 0; k <= SECTIONS; ++k) {       // rotate_push_back on every line, after all reads
            s[2 * k] = s[2 * k + 1];
            s[2 * k + 1] = sig[k];
        }
    }
};

}  // namespace zgk
