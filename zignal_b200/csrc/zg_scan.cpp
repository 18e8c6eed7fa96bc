// zignal-b200 :: what a time-parallel evaluation of a linear tick has to know (host analysis, float64).
//
// One tick of a LINEAR / AFFINE graph is  state' = A state + B x + c,  y = C state + D x + d  (literals and $k
// parameters are coefficients).  The reference evaluates it one sample after the other, one voice per object
// (binary_feedback, flowz/flowz.hpp:1031-1074: "eval future -> eval promise -> push"); the state-space reading is
// sketched in experimental_steps/tuprix.cpp:239-254.  With few, long channels the device kernels cut time into
// segments (kernels/zg_stream.cuh, StreamArgs::n_segs) and need, per channel,
//   * A itself, read off the tick program by unit-vector probes of the host tick (every entry is one coefficient path),
//   * how fast the tick forgets its state: the smallest K with |A^K|_inf <= tol (warm-up form), and
//   * A^L for the boundary fix-up x <- A^L x + z of the two-pass form.
#include <algorithm>
#include <cmath>

#include "zg_internal.hpp"

namespace zg {

int ir_linearity(const Ir& ir) {
    enum { CONST = 0, LIN = 1, AFF = 2, NONLIN = 3 };
    std::vector<int> cls(ir.nodes.size(), CONST);
    for (size_t i = 0; i < ir.nodes.size(); ++i) {
        const IrNode& n = ir.nodes[i];
        switch (n.op) {
            case IrOp::In: case IrOp::DRead: cls[i] = LIN; break;
            case IrOp::Const: case IrOp::Param: cls[i] = CONST; break;
            case IrOp::Neg: cls[i] = cls[n.a]; break;
            case IrOp::Add: case IrOp::Sub: {
                const int a = cls[n.a], b = cls[n.b];
                cls[i] = (a == NONLIN || b == NONLIN) ? NONLIN : a == b && a != AFF ? a : (a == CONST && b == CONST) ? CONST : AFF;
                break;
            }
            case IrOp::Mul: {
                const int a = cls[n.a], b = cls[n.b];
                cls[i] = a == CONST ? b : b == CONST ? a : NONLIN;
                break;
            }
            case IrOp::Div: cls[i] = cls[n.b] == CONST ? cls[n.a] : NONLIN; break;
            default: cls[i] = NONLIN; break;
        }
    }
    int worst = LIN;
    auto see = [&](int id) {
        const int c = cls[id] == CONST ? (ir.nodes[id].op == IrOp::Const && ir.nodes[id].value == 0 ? LIN : AFF) : cls[id];
        worst = std::max(worst, c);
    };
    for (int o : ir.outs) see(o);
    for (const IrLine& l : ir.lines) see(l.src);
    return worst == NONLIN ? ZG_NONLINEAR : worst == AFF ? ZG_AFFINE : ZG_LINEAR;
}

bool tick_matrix(const Ir& ir, const float* params, std::vector<double>& A) {
    const int n = ir.n_state;
    A.assign((size_t)n * n, 0.0);
    std::vector<float> st(n), base(n);
    std::vector<double> in(std::max(ir.n_in, 1), 0.0), out(std::max(ir.n_out, 1));
    host_tick(ir, base.data(), params, in.data(), out.data());          // the constant term c of an affine tick
    for (int j = 0; j < n; ++j) {
        std::fill(st.begin(), st.end(), 0.f);
        st[j] = 1.f;
        host_tick(ir, st.data(), params, in.data(), out.data());
        for (int i = 0; i < n; ++i) {
            const double v = (double)st[i] - (double)base[i];
            if (!std::isfinite(v)) return false;
            A[(size_t)i * n + j] = v;
        }
    }
    return true;
}

static void mat_mul(const std::vector<double>& X, const std::vector<double>& Y, int n, std::vector<double>& Z) {
    Z.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            const double x = X[(size_t)i * n + k];
            if (x == 0.0) continue;
            for (int j = 0; j < n; ++j) Z[(size_t)i * n + j] += x * Y[(size_t)k * n + j];
        }
}

void mat_pow(const std::vector<double>& A, int n, long e, std::vector<double>& out) {
    std::vector<double> base = A, tmp;
    out.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) out[(size_t)i * n + i] = 1.0;
    while (e > 0) {
        if (e & 1) { mat_mul(out, base, n, tmp); out.swap(tmp); }
        e >>= 1;
        if (e) { mat_mul(base, base, n, tmp); base.swap(tmp); }
    }
}

static double norm_inf(const std::vector<double>& M, int n) {
    double worst = 0;
    for (int i = 0; i < n; ++i) {
        double row = 0;
        for (int j = 0; j < n; ++j) row += std::fabs(M[(size_t)i * n + j]);
        if (!(row <= worst)) worst = row;            // NaN propagates
    }
    return worst;
}

int decay_length(const std::vector<double>& A, int n, int step, int k_max, double tol) {
    if (n == 0) return step;
    std::vector<double> P, Q, tmp;
    mat_pow(A, n, step, P);
    Q = P;
    for (int K = step; K <= k_max; K += step) {
        const double nq = norm_inf(Q, n);
        if (!std::isfinite(nq)) return 0;
        if (nq <= tol) return K;
        if (nq > 1e30) return 0;                     // growing: an unstable tick never forgets
        mat_mul(Q, P, n, tmp);
        Q.swap(tmp);
    }
    return 0;
}

}  // namespace zg
