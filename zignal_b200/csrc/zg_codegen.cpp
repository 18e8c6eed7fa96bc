// zignal-b200 :: tick program -> CUDA tick functor (the generated part of the generic kernel, K2).
//
// The hand-written part is kernels/zg_stream.cuh; this file only prints the straight-line body of
// one tick in SSA order.  It stands in for what the reference gets from template expansion of
// eval_it (flowz/flowz.hpp:740-774) over the canonical expression.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <sstream>

#include "zg_internal.hpp"

namespace zg {

namespace {

const char* ctype(Dtype d) { return d == Dtype::I32 ? "int" : d == Dtype::F32 ? "float" : "double"; }

std::string literal(Dtype d, double v) {
    char buf[64];
    if (d == Dtype::I32) std::snprintf(buf, sizeof buf, "%d", (int)v);
    else if (d == Dtype::F32) {
        if (v != v) return "__int_as_float(0x7fc00000)";
        if (v == 1.0 / 0.0) return "__int_as_float(0x7f800000)";
        if (v == -1.0 / 0.0) return "__int_as_float(0xff800000)";
        std::snprintf(buf, sizeof buf, "%af", v);
    } else {
        std::snprintf(buf, sizeof buf, "%a", v);
    }
    return buf;
}

}  // namespace

// Delayed-product reuse (EXACT mode).  `c * line[n]` -- a coefficient times the value a delay line received n ticks
// ago -- is the product `c * src` that this very tick computes for the line's current input, n ticks late: the same
// two operands, hence the same correctly rounded bits.  When both appear in one tick (b0*x and b2*x[-2] of every RBJ
// low-pass / high-pass / notch section, where b0 == b2; symmetric FIR taps; ...) the product is carried in a short
// register line of its own instead of being recomputed: one instruction less per match per sample, bit-identical.
// The carried products are not delay-line state: they are rebuilt from the state at block start (init), so the state
// that crosses the ABI keeps the reference's layout.  (FAST mode folds the product into an FMA: nothing to save.)
struct ProductLine {
    int product;        // node computing c * src in this tick
    int coef;           // the coefficient node (Const or Param)
    int line;           // the delay line whose input is src
    int depth = 0;      // deepest delayed use
    int base = 0;       // first extra register
};

static bool same_coefficient(const Ir& ir, int a, int b) {
    const IrNode &x = ir.nodes[a], &y = ir.nodes[b];
    if (x.op != y.op || x.dtype != Dtype::F32 || y.dtype != Dtype::F32) return false;
    if (x.op == IrOp::Param) return x.a == y.a;
    if (x.op != IrOp::Const) return false;
    const float u = (float)x.value, w = (float)y.value;
    return std::memcmp(&u, &w, sizeof u) == 0;
}

// reuse[m] = (product line index, n) for every Mul node m that becomes a read of a carried product
static std::vector<ProductLine> find_product_reuse(const Ir& ir, std::vector<std::pair<int, int>>& reuse) {
    std::vector<ProductLine> out;
    reuse.assign(ir.nodes.size(), {-1, 0});
    auto is_coef = [&](int id) { return ir.nodes[id].dtype == Dtype::F32 && (ir.nodes[id].op == IrOp::Const || ir.nodes[id].op == IrOp::Param); };
    for (size_t m = 0; m < ir.nodes.size(); ++m) {
        const IrNode& n = ir.nodes[m];
        if (n.op != IrOp::Mul || n.dtype != Dtype::F32) continue;
        int c = -1, r = -1;
        if (is_coef(n.a) && ir.nodes[n.b].op == IrOp::DRead) { c = n.a; r = n.b; }
        else if (is_coef(n.b) && ir.nodes[n.a].op == IrOp::DRead) { c = n.b; r = n.a; }
        if (c < 0 || ir.nodes[r].dtype != Dtype::F32) continue;
        const int line = ir.nodes[r].a, delay = ir.nodes[r].b;
        const int src = ir.lines[line].src;
        if (ir.nodes[src].dtype != Dtype::F32) continue;         // the push narrows to float (flowz.hpp:136)
        // the same coefficient times the line's input, computed in this tick
        for (size_t q = 0; q < ir.nodes.size(); ++q) {
            const IrNode& k = ir.nodes[q];
            if (k.op != IrOp::Mul || k.dtype != Dtype::F32) continue;
            const int kc = k.a == src ? k.b : k.b == src ? k.a : -1;
            if (kc < 0 || !is_coef(kc) || !same_coefficient(ir, kc, c)) continue;
            int idx = -1;
            for (size_t i = 0; i < out.size(); ++i)
                if (out[i].product == (int)q && out[i].line == line) idx = (int)i;
            if (idx < 0) {
                idx = (int)out.size();
                ProductLine pl;
                pl.product = (int)q;
                pl.coef = kc;
                pl.line = line;
                out.push_back(pl);
            }
            out[idx].depth = std::max(out[idx].depth, delay);
            reuse[m] = {idx, delay};
            break;
        }
    }
    int base = 0;
    for (ProductLine& pl : out) { pl.base = base; base += pl.depth; }
    return out;
}

std::string generate_tick_source(const Ir& ir, bool exact, const std::string& struct_name, int n_ring_in,
                                 int n_ring_out, int ring_pf) {
    std::ostringstream os;
    std::vector<std::pair<int, int>> reuse(ir.nodes.size(), {-1, 0});
    std::vector<ProductLine> products;
    if (exact && !std::getenv("ZG_TUNE_NO_PRODUCT_REUSE")) products = find_product_reuse(ir, reuse);
    int n_extra = 0;
    for (const ProductLine& pl : products) n_extra += pl.depth;
    auto ref = [&](int id, Dtype to) {
        std::ostringstream r;
        if (ir.nodes[id].dtype == to) r << "v" << id;
        else r << "(" << ctype(to) << ")v" << id;
        return r.str();
    };
    os << "struct " << struct_name << " {\n";
    os << "    static constexpr int N_IN = " << ir.n_in << ", N_OUT = " << ir.n_out << ", N_STATE = " << ir.n_state
       << ", N_PARAM = " << ir.n_params << ";\n";
    os << "    static constexpr unsigned SYNTH_MASK = ZG_SYNTH_MASK;\n";
    // the last n_ring_in inputs / n_ring_out outputs are far reads / pushes of long delay lines (zg_ir.hpp)
    if (n_ring_in || n_ring_out)
        os << "    static constexpr int N_RING_IN = " << n_ring_in << ", N_RING_OUT = " << n_ring_out
           << ", RING_PF = " << ring_pf << ";\n";
    {
        // keep the unrolled loop body of the skeleton around 600 instructions (kernels/zg_stream.cuh)
        int arith = 0;
        for (const IrNode& n : ir.nodes)
            if (n.op == IrOp::Add || n.op == IrOp::Sub || n.op == IrOp::Mul || n.op == IrOp::Div || n.op == IrOp::Neg) ++arith;
        int cu = 8;
        while (cu > 1 && arith * 4 * cu > 640) cu /= 2;
        if (const char* e = std::getenv("ZG_TUNE_CHUNK_UNROLL")) { int v = std::atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) cu = v; }
        os << "    static constexpr int CHUNK_UNROLL = " << cu << ";\n";
    }
    if (n_extra > 0) {
        // carried products: e[base + depth - n] = coefficient * (value the line received n ticks ago)
        os << "    static constexpr int N_EXTRA = " << n_extra << ";\n";
        os << "    template <class P>\n"
              "    static __device__ __forceinline__ void init(const zgk::Arr<N_STATE>& s, const P& p, zgk::Arr<N_EXTRA>& e) {\n";
        for (const ProductLine& pl : products) {
            const IrNode& c = ir.nodes[pl.coef];
            const IrLine& l = ir.lines[pl.line];
            std::string coef = c.op == IrOp::Param ? "p[" + std::to_string(c.a) + "]" : literal(Dtype::F32, c.value);
            for (int j = 1; j <= pl.depth; ++j)
                os << "        e[" << pl.base + pl.depth - j << "] = __fmul_rn(" << coef << ", s[" << l.offset + l.depth - j << "]);\n";
        }
        os << "    }\n";
    }
    os << "    template <class P>\n"
          "    static __device__ __forceinline__ void tick(const zgk::Arr<N_IN>& x, zgk::Arr<N_OUT>& y,\n"
          "                                                zgk::Arr<N_STATE>& s, const P& p"
       << (n_extra > 0 ? ", zgk::Arr<N_EXTRA>& e" : "") << ") {\n";
    for (size_t i = 0; i < ir.nodes.size(); ++i) {
        const IrNode& n = ir.nodes[i];
        const char* t = ctype(n.dtype);
        os << "        const " << t << " v" << i << " = ";
        if (reuse[i].first >= 0) {
            const ProductLine& pl = products[reuse[i].first];
            os << "e[" << pl.base + pl.depth - reuse[i].second << "];\n";
            continue;
        }
        switch (n.op) {
            case IrOp::In: os << "(" << t << ")x[" << n.a << "]"; break;
            case IrOp::Const: os << literal(n.dtype, n.value); break;
            case IrOp::Param: os << "p[" << n.a << "]"; break;
            case IrOp::DRead: os << "s[" << (ir.lines[n.a].offset + ir.lines[n.a].depth - n.b) << "]"; break;
            case IrOp::Neg: os << "-v" << n.a; break;
            default: {
                std::string a = ref(n.a, n.dtype), b = ref(n.b, n.dtype);
                const char* sym = n.op == IrOp::Add ? "+" : n.op == IrOp::Sub ? "-" : n.op == IrOp::Mul ? "*" : "/";
                if (n.dtype == Dtype::I32) {
                    if (n.op == IrOp::Div) os << "(" << b << " == 0 ? 0 : " << b << " == -1 ? (int)(0u - (unsigned)" << a << ") : " << a << " / " << b << ")";
                    else os << a << " " << sym << " " << b;
                } else if (exact) {
                    // separately rounded, never contracted into FMA
                    const char* pre = n.dtype == Dtype::F32 ? "__f" : "__d";
                    const char* nm = n.op == IrOp::Add ? "add" : n.op == IrOp::Sub ? "sub" : n.op == IrOp::Mul ? "mul" : "div";
                    os << pre << nm << "_rn(" << a << ", " << b << ")";
                } else {
                    os << a << " " << sym << " " << b;
                }
            }
        }
        os << ";\n";
    }
    for (int o = 0; o < ir.n_out; ++o) os << "        y[" << o << "] = (float)v" << ir.outs[o] << ";\n";
    // rotate_push_back for every line; all reads above already happened
    for (const IrLine& l : ir.lines) {
        for (int j = 0; j + 1 < l.depth; ++j) os << "        s[" << l.offset + j << "] = s[" << l.offset + j + 1 << "];\n";
        os << "        s[" << l.offset + l.depth - 1 << "] = (float)v" << l.src << ";\n";
    }
    for (const ProductLine& pl : products) {
        for (int j = 0; j + 1 < pl.depth; ++j) os << "        e[" << pl.base + j << "] = e[" << pl.base + j + 1 << "];\n";
        os << "        e[" << pl.base + pl.depth - 1 << "] = v" << pl.product << ";\n";
    }
    os << "    }\n};\n";
    return os.str();
}

}  // namespace zg
