// zignal-b200 :: recognisers that route a tick program to a prebuilt kernel.
//
// match_df1_cascade(): is the tick program a series of direct-form-1 biquads exactly as the
// reference spells them (test/benchmark.cpp:25-33, `fwd |= bwd` chained with |=)?  The match is on
// the lowered SSA, so any spelling that lowers to the same arithmetic in the same association is
// accepted (|= or >>, literal or std::ref coefficients) and anything else falls through to the
// generated kernel -- never to a different arithmetic.
#include "zg_internal.hpp"

namespace zg {

namespace {

bool is_coef(const Ir& ir, int id) {
    const IrNode& n = ir.nodes[id];
    return n.dtype == Dtype::F32 && (n.op == IrOp::Const || n.op == IrOp::Param);
}

BiquadCoef coef_of(const Ir& ir, int id) {
    const IrNode& n = ir.nodes[id];
    BiquadCoef c;
    c.is_param = n.op == IrOp::Param;
    c.param = c.is_param ? n.a : -1;
    c.value = c.is_param ? 0.f : (float)n.value;
    return c;
}

// node == coef * other  (either operand order; fp32 multiplication commutes bit for bit)
bool match_scaled(const Ir& ir, int id, BiquadCoef& c, int& other) {
    const IrNode& n = ir.nodes[id];
    if (n.op != IrOp::Mul || n.dtype != Dtype::F32) return false;
    if (is_coef(ir, n.a) && !is_coef(ir, n.b)) { c = coef_of(ir, n.a); other = n.b; return true; }
    if (is_coef(ir, n.b) && !is_coef(ir, n.a)) { c = coef_of(ir, n.b); other = n.a; return true; }
    return false;
}

bool match_dread(const Ir& ir, int id, int& line, int n) {
    const IrNode& d = ir.nodes[id];
    if (d.op != IrOp::DRead || d.b != n) return false;
    line = d.a;
    return true;
}

// node == (acc + c1 * line[-1]) + c2 * line[-2]
bool match_two_taps(const Ir& ir, int id, int& acc, BiquadCoef& c1, BiquadCoef& c2, int& line) {
    const IrNode& top = ir.nodes[id];
    if (top.op != IrOp::Add || top.dtype != Dtype::F32) return false;
    int d2, l2;
    if (!match_scaled(ir, top.b, c2, d2) || !match_dread(ir, d2, l2, 2)) return false;
    const IrNode& mid = ir.nodes[top.a];
    if (mid.op != IrOp::Add || mid.dtype != Dtype::F32) return false;
    int d1, l1;
    if (!match_scaled(ir, mid.b, c1, d1) || !match_dread(ir, d1, l1, 1)) return false;
    if (l1 != l2) return false;
    acc = mid.a;
    line = l1;
    return true;
}

}  // namespace

bool match_df1_cascade(const Ir& ir, BiquadMatch& m) {
    if (ir.n_in != 1 || ir.n_out != 1 || !ir.all_f32()) return false;
    std::vector<std::array<BiquadCoef, 5>> rev;
    std::vector<int> rev_line;       // line of the section's output, last section first
    int y = ir.outs[0];
    int in_line = -1;
    for (;;) {
        if (ir.nodes[y].op == IrOp::In) break;
        if ((int)rev.size() == kMaxBiquadSections) return false;
        // y = (v + a1*y1) + a2*y2
        int v, ly;
        BiquadCoef a1, a2;
        if (!match_two_taps(ir, y, v, a1, a2, ly)) return false;
        if (ir.lines[ly].src != y || ir.lines[ly].depth != 2) return false;
        // v = (b0*x + b1*x1) + b2*x2
        int head, lx;
        BiquadCoef b0, b1, b2;
        if (!match_two_taps(ir, v, head, b1, b2, lx)) return false;
        int x;
        if (!match_scaled(ir, head, b0, x)) return false;
        if (ir.lines[lx].src != x || ir.lines[lx].depth != 2) return false;
        rev.push_back({b0, b1, b2, a1, a2});
        rev_line.push_back(ly);
        in_line = lx;
        y = x;
    }
    const int S = (int)rev.size();
    if (S == 0 || ir.nodes[y].a != 0) return false;
    if ((int)ir.lines.size() != S + 1 || ir.n_state != 2 * (S + 1)) return false;
    m.sections = S;
    m.signal_line.assign(S + 1, -1);
    m.signal_line[0] = in_line;
    for (int k = 0; k < S; ++k) {
        m.coef[k] = rev[S - 1 - k];
        m.signal_line[k + 1] = rev_line[S - 1 - k];
    }
    // section k's input line must be section k-1's output line (shared after line merging)
    // -- re-walk to verify the chain is consistent
    {
        int yy = ir.outs[0];
        for (int k = S - 1; k >= 0; --k) {
            int v, ly, head, lx, x;
            BiquadCoef t1, t2, t0;
            match_two_taps(ir, yy, v, t1, t2, ly);
            match_two_taps(ir, v, head, t1, t2, lx);
            match_scaled(ir, head, t0, x);
            if (ly != m.signal_line[k + 1] || lx != m.signal_line[k]) return false;
            yy = x;
        }
    }
    // every line exactly once
    std::vector<char> seen(ir.lines.size(), 0);
    for (int l : m.signal_line) {
        if (l < 0 || seen[l]) return false;
        seen[l] = 1;
    }
    return true;
}

// match_fir(): a single-input, single-output sum of N >= 2 terms coef * x[-k], k = 0..N-1 in that
// order, associated to the left.  The kernel (kernels/zg_fir.cuh) adds the products in exactly this
// order, so EXACT mode stays bit-identical; any other spelling (sparse taps, another order, a
// bracketed sub-sum) is NOT re-associated to fit -- it falls through to the generated kernel.
bool match_fir(const Ir& ir, FirMatch& m) {
    if (ir.n_in != 1 || ir.n_out != 1 || !ir.all_f32() || ir.lines.size() != 1) return false;
    const IrLine& line = ir.lines[0];
    if (line.src < 0 || ir.nodes[line.src].op != IrOp::In || ir.nodes[line.src].a != 0) return false;
    std::vector<std::pair<int, BiquadCoef>> rev;      // (delay, coef), last term first
    int id = ir.outs[0];
    auto term = [&](int t, int& delay, BiquadCoef& c) {
        int other;
        if (!match_scaled(ir, t, c, other)) return false;
        const IrNode& o = ir.nodes[other];
        if (o.op == IrOp::In && o.a == 0) { delay = 0; return true; }
        if (o.op == IrOp::DRead && o.a == 0) { delay = o.b; return true; }
        return false;
    };
    for (;;) {
        const IrNode& n = ir.nodes[id];
        int delay;
        BiquadCoef c;
        if (n.op == IrOp::Add && n.dtype == Dtype::F32) {
            if (!term(n.b, delay, c)) return false;
            rev.push_back({delay, c});
            id = n.a;
            continue;
        }
        if (!term(id, delay, c)) return false;
        rev.push_back({delay, c});
        break;
    }
    const int N = (int)rev.size();
    if (N < 2 || line.depth != N - 1 || ir.n_state != N - 1) return false;
    m.taps.resize(N);
    for (int k = 0; k < N; ++k) {
        if (rev[N - 1 - k].first != k) return false;
        m.taps[k] = rev[N - 1 - k].second;
    }
    return true;
}

}  // namespace zg
