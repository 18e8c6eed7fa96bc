// oracle/_ref driver -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference's own benchmark translation unit (test/benchmark.cpp) *where it lies* under
// $(REF) -- no reference source is copied into this repository -- and exports, behind a C ABI,
//   * the reference's hand-written biquad loops (`make_custom`, test/benchmark.cpp:35-47, 65-76,
//     90-105, 116-126): the Boost-free yardstick the reference measures flowz against.  These pin
//     the oracle's biquad numerics (tests/test_oracle_ref.py) and are the timed CPU baseline of
//     `bench.py --impl reference` (cpu_baseline.kind = "reference");
//   * the reference's flowz graphs (`make_flow`, :29-32, 60-63, 83-86, 111-114) compiled against
//     THIS repository's include/flowz/flowz.hpp shim -- i.e. proof that the unmodified reference
//     source builds and runs on the shim (drop-in check).
// <benchmark/benchmark.h> resolves to oracle/stubs; <flowz/flowz.hpp> resolves to include/.
//
// Built twice (oracle/Makefile) so that the half that pins the oracle and times the CPU arm cannot touch the product:
//   _ref/libzg_ref_custom.so  -DZG_REF_PART=1: zg_ref_custom, zg_ref_sum_dirac_custom, zg_ref_df1_chain -- the reference's
//                             own loops only; linked with --no-undefined and WITHOUT libzignal_b200;
//   _ref/libzg_ref_flow.so    -DZG_REF_PART=2: zg_ref_flow, zg_ref_sum_dirac_flow -- the reference's make_flow() graphs on
//                             this repository's shim, linked against libzignal_b200 (the drop-in check).
#include <test/benchmark.cpp>

#ifndef ZG_REF_PART
#error "compile with -DZG_REF_PART=1 (reference loops only) or -DZG_REF_PART=2 (make_flow on the shim)"
#endif

#include <cstddef>
#include <tuple>

namespace {

template <class F>
void run_block(F f, const float* x, float* y, long n) {
    for (long t = 0; t < n; ++t) y[t] = std::get<0>(f(x[t]));
}

}  // namespace

extern "C" {

#if ZG_REF_PART == 1
// form: 1 = DF1, 2 = DF2, 3 = DF1 transposed, 4 = DF2 transposed.  Fresh (zero) state, one voice.
int zg_ref_custom(int form, const float* x, float* y, long n) {
    switch (form) {
        case 1: run_block(biquad::direct_form_1::make_custom(), x, y, n); return 0;
        case 2: run_block(biquad::direct_form_2::make_custom(), x, y, n); return 0;
        case 3: run_block(biquad::direct_form_1_transposed::make_custom(), x, y, n); return 0;
        case 4: run_block(biquad::direct_form_2_transposed::make_custom(), x, y, n); return 0;
    }
    return -1;
}

#endif
#if ZG_REF_PART == 2
// the same four graphs as flowz expressions, ticked through the shim
int zg_ref_flow(int form, const float* x, float* y, long n) {
    try {
        switch (form) {
            case 1: run_block(biquad::direct_form_1::make_flow(), x, y, n); return 0;
            case 2: run_block(biquad::direct_form_2::make_flow(), x, y, n); return 0;
            case 3: run_block(biquad::direct_form_1_transposed::make_flow(), x, y, n); return 0;
            case 4: run_block(biquad::direct_form_2_transposed::make_flow(), x, y, n); return 0;
        }
    } catch (...) {
    }
    return -1;
}

#endif
#if ZG_REF_PART == 1
// sum_dirac (test/benchmark.cpp:137-147) on the custom loop / on the flowz graph
float zg_ref_sum_dirac_custom(int form) {
    switch (form) {
        case 1: { auto f = biquad::direct_form_1::make_custom(); return sum_dirac(f); }
        case 2: { auto f = biquad::direct_form_2::make_custom(); return sum_dirac(f); }
        case 3: { auto f = biquad::direct_form_1_transposed::make_custom(); return sum_dirac(f); }
        case 4: { auto f = biquad::direct_form_2_transposed::make_custom(); return sum_dirac(f); }
    }
    return 0.f;
}
#endif
#if ZG_REF_PART == 2
float zg_ref_sum_dirac_flow(int form) {
    switch (form) {
        case 1: { auto f = biquad::direct_form_1::make_flow(); return sum_dirac(f); }
        case 2: { auto f = biquad::direct_form_2::make_flow(); return sum_dirac(f); }
        case 3: { auto f = biquad::direct_form_1_transposed::make_flow(); return sum_dirac(f); }
        case 4: { auto f = biquad::direct_form_2_transposed::make_flow(); return sum_dirac(f); }
    }
    return 0.f;
}

#endif
#if ZG_REF_PART == 1
// The reference's own CPU loop for the benchmark workload: `sections` hand-written DF1 biquads in
// series per channel (the way make_custom2, test/benchmark.cpp:49-55, chains two), one sample per
// call, channels spread over the host cores.  Planar [channels][n].  Coefficients are the
// reference's compile-time constants (:18-23).
int zg_ref_df1_chain(int sections, const float* x, float* y, long channels, long n) {
    if (sections < 1 || sections > 8) return -1;
    #pragma omp parallel for schedule(static)
    for (long c = 0; c < channels; ++c) {
        using F = decltype(biquad::direct_form_1::make_custom());
        F f[8] = {biquad::direct_form_1::make_custom(), biquad::direct_form_1::make_custom(),
                  biquad::direct_form_1::make_custom(), biquad::direct_form_1::make_custom(),
                  biquad::direct_form_1::make_custom(), biquad::direct_form_1::make_custom(),
                  biquad::direct_form_1::make_custom(), biquad::direct_form_1::make_custom()};
        const float* xc = x + c * n;
        float* yc = y + c * n;
        for (long t = 0; t < n; ++t) {
            float v = xc[t];
            for (int s = 0; s < sections; ++s) v = std::get<0>(f[s](v));
            yc[t] = v;
        }
    }
    return 0;
}
#endif

}  // extern "C"
