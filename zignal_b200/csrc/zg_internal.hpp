// zignal-b200 :: objects behind the opaque C handles
#pragma once
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zignal_b200.h"
#include "zg_ir.hpp"

struct zg_graph {
    std::string text;
    zg::ExprP user;
    zg::ExprP canonical;
    int n_in = 0, n_out = 0;
    zg::Ir ir_f32;                 // the tick program for fp32 inputs (device path + float ticks)
    std::string canonical_str, dump_str;
    mutable std::mutex mu;
    mutable std::map<std::vector<zg::Dtype>, std::shared_ptr<const zg::Ir>> irs;

    std::shared_ptr<const zg::Ir> ir_for(const std::vector<zg::Dtype>& sig) const;
};

struct zg_voice {
    const zg_graph* g;
    std::vector<float> state;
    std::vector<float> params;
};

namespace zg {
void set_last_error(const std::string& msg);
int fail(int status, const std::string& msg);

// CUDA tick-functor source for one IR (zg_codegen.cpp)
std::string generate_tick_source(const Ir& ir, bool exact, const std::string& struct_name, int n_ring_in = 0,
                                 int n_ring_out = 0, int ring_pf = 1);

// ---- linear ticks (zg_scan.cpp): what the time-segmented launches of few, long channels need ----
int ir_linearity(const Ir& ir);                                   // zg_linearity of one tick
// A of  state' = A state + B x + c  (row-major [n_state][n_state], float64) for one set of parameter values;
// false when an entry is not finite
bool tick_matrix(const Ir& ir, const float* params, std::vector<double>& A);
void mat_pow(const std::vector<double>& A, int n, long e, std::vector<double>& out);
// smallest K = m * step <= k_max with |A^K|_inf <= tol, 0 if there is none (a tick that does not forget its state)
int decay_length(const std::vector<double>& A, int n, int step, int k_max, double tol);

// ---- prebuilt-kernel recognisers (zg_match.cpp) ----
constexpr int kMaxBiquadSections = 8;
struct BiquadCoef {
    bool is_param = false;
    int param = -1;      // $index when is_param
    float value = 0.f;   // literal otherwise
};
struct BiquadMatch {
    int sections = 0;
    std::array<BiquadCoef, 5> coef[kMaxBiquadSections];   // b0 b1 b2 a1 a2 per section
    std::vector<int> signal_line;                         // line of signal k (0 = input, k = out of section k)
};
bool match_df1_cascade(const Ir& ir, BiquadMatch& m);

// out = ((c0*x + c1*x[-1]) + c2*x[-2]) + ... + c(N-1)*x[-(N-1)]: a dense FIR summed left to right, the
// association of the C++ parse tree of `c0*_1 + c1*_1[_1] + ...` (BASELINE configs[3])
struct FirMatch {
    std::vector<BiquadCoef> taps;    // taps[k] multiplies the input delayed by k samples
};
bool match_fir(const Ir& ir, FirMatch& m);
}  // namespace zg
