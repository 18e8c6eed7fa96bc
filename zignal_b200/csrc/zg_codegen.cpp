// zignal-b200 :: tick program -> CUDA tick functor (the generated part of the generic kernel, K2).
//
// The hand-written part is kernels/zg_stream.cuh; this file only prints the straight-line body of
// one tick in SSA order.  It stands in for what the reference gets from template expansion of
// eval_it (flowz/flowz.hpp:740-774) over the canonical expression.
#include <cstdio>
#include <cstdlib>
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

std::string generate_tick_source(const Ir& ir, bool exact, const std::string& struct_name, int n_ring_in,
                                 int n_ring_out, int ring_pf) {
    std::ostringstream os;
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
    os << "    template <class P>\n"
          "    static __device__ __forceinline__ void tick(const zgk::Arr<N_IN>& x, zgk::Arr<N_OUT>& y,\n"
          "                                                zgk::Arr<N_STATE>& s, const P& p) {\n";
    for (size_t i = 0; i < ir.nodes.size(); ++i) {
        const IrNode& n = ir.nodes[i];
        const char* t = ctype(n.dtype);
        os << "        const " << t << " v" << i << " = ";
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
    os << "    }\n};\n";
    return os.str();
}

}  // namespace zg
