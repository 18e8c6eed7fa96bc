// zignal-b200 :: flat tick program (IR), lowering from the canonical expression, host interpreter.
//
// One tick of a canonical flowz expression is a pure function
//      (delay lines, inputs) -> (outputs, one pushed value per delay line)
// because within a tick every line is read before it is pushed and has exactly one writer
// (reference: sequence pushes at flowz/flowz.hpp:994, binary_feedback at :1067, both after the
// sub-expressions that read the line have been evaluated).  lower() walks the canonical tree once
// with *symbolic* wire values, following the routing rules of the reference evaluators
// (sequence :960-1001, binary_feedback :1031-1074, parallel :1076-1101, channel :765-768,
// place_the_holder/place_delay :941-958) and records the result as straight-line SSA.
// A unary feedback that canonical_with_front() kept whole (a graph the reference cannot compile) is walked with forward
// references for its fed-back wires; they are bound when the loop closes and the SSA is re-sorted topologically, so the
// result is the same kind of program -- or an Error if a loop has no delay in it.
#pragma once
#include "zg_expr.hpp"

namespace zg {

enum class IrOp : uint8_t { In, Const, Param, DRead, Neg, Add, Sub, Mul, Div,
                            Fwd /* lower() only: current value of a fed-back wire, bound when the loop closes */ };

struct IrNode {
    IrOp op;
    Dtype dtype;
    int a = -1, b = -1;   // operands (node ids) | In: a = input index | Param: a = param index
                          // DRead: a = line, b = n  (value pushed n ticks ago)
    double value = 0;     // Const
};

struct IrLine {
    int depth = 0;        // number of floats kept; slot depth-n holds the value pushed n ticks ago
    int offset = 0;       // first float of this line inside the per-channel state vector
    int src = -1;         // node pushed at the end of every tick (narrowed to float, :136)
};

struct Ir {
    int n_in = 0, n_out = 0, n_params = 0, n_state = 0;
    std::vector<Dtype> in_dtypes;
    std::vector<IrNode> nodes;     // topologically ordered
    std::vector<IrLine> lines;
    std::vector<int> outs;         // node ids
    bool all_f32() const;
    std::string dump() const;      // text form, stands in for the reference's print_state/demangle
};

struct LowerOptions {
    bool merge_lines = true;       // share delay lines that are fed by the same node
                                   // (removes the duplicates the reference keeps, TODO.md:33-34,59)
    bool cse = true;               // common sub-expression elimination (bit-exact, same ops)
};

// `canonical` must come from canonical_with_front(); in_dtypes holds one dtype per input of the USER expression.
Ir lower(const Expr& canonical, const std::vector<Dtype>& in_dtypes, const LowerOptions& opt = {});

// ---- long delay lines on the device (zg_ir.cpp: split_long_lines) -----------------------------------------
// A delay line deeper than `reg_depth` floats does not live in registers.  Its D floats of state stay where the
// state layout puts them, [D rows][channels] in HBM, used as a ring: the value pushed at absolute tick t sits
// in row t mod D.  The kernel-side tick program is the same program with
//   * every far read  DRead(line, n >= far)   turned into an extra INPUT  (the skeleton loads ring row
//     (t - n) mod D a chunk ahead of use: a coalesced 128-byte load per warp),
//   * the line's pushed value                 turned into an extra OUTPUT (the skeleton stores it to row t mod D),
//   * near reads      DRead(line, n <  far)   served by a short register window of the last `near` pushes
//     (reloaded from the ring at block start; never written back: the ring already holds those values).
struct RingTap { int line; int n; };           // extra kernel input k <-> original line, delay
struct KernelSlot {                            // kernel register-state slot -> where it lives in d_state
    int row = 0;                               // fixed row (short lines), or first row of the ring (windows)
    int ring_depth = 0;                        // 0: fixed row; else the slot holds the value pushed `ago` ticks ago
    int ago = 0;
};
struct RingPlan {
    std::vector<RingTap> taps;                 // in kernel-input order, after the graph's own inputs
    std::vector<int> out_lines;                // original line of extra kernel output k, after the graph's outputs
    std::vector<KernelSlot> slots;             // one per kernel state slot (kir.n_state)
    bool any() const { return !out_lines.empty(); }
};
// Returns the kernel-side program.  Lines with depth <= reg_depth are untouched (their slots map to fixed rows).
Ir split_long_lines(const Ir& ir, int reg_depth, int far, RingPlan& rp);

// Host scalar tick: this is stateful_lambda::operator() (flowz/flowz.hpp:1193-1201, 1225-1229)
// for one voice.  `state` has ir.n_state floats, zero-initialised by the caller (:1191 value-init).
// `params` has ir.n_params floats.  Inputs/outputs travel as doubles and are converted to/from
// the node dtype (an int input stays an int inside the tick, as in the reference).
void host_tick(const Ir& ir, float* state, const float* params, const double* in, double* out);

// Block version of the host tick for one channel, fp32 in/out (used by the C++ shim's
// process_host(); the device path is zg_process()).
void host_block_f32(const Ir& ir, float* state, const float* params, const float* const* in,
                    float* const* out, long n_samples, long in_stride, long out_stride);

}  // namespace zg
