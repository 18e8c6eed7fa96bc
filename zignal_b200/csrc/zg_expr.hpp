// zignal-b200 :: expression layer (host, C++17, no Boost)
//
// Run-time restatement of the flowz expression grammar and of its static analyses.  The reference
// does all of this with Boost.Proto at C++ compile time; here the tree is a run-time value so that
// the same code serves the C ABI, the C++ EDSL shim (include/flowz/flowz.hpp) and the Python tests.
//
// Reference behaviour mirrored (all paths relative to /root/reference):
//   grammar / building blocks        flowz/flowz.hpp:68-102   (placeholders, _k[_n], |= | , ~, bfb)
//   input_arity / output_arity       flowz/flowz.hpp:162-246
//   max/min_input_delays             flowz/flowz.hpp:286-506
//   make_front / add_front_panel     flowz/flowz.hpp:261-277
//   make_canonical and friends       flowz/flowz.hpp:794-935
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace zg {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// C64 / C128 (std::complex<float> / <double>) exist for the type analysis only (result_types): a graph with a
// complex terminal is analysed like any other but cannot be lowered to a tick program.
enum class Dtype : uint8_t { I32 = 0, F32 = 1, F64 = 2, C64 = 4, C128 = 5 };
// Every analysis walks the tree recursively: the height is bounded so that no expression text can exhaust the stack
// (a 256-tap FIR written as one sum is 256 levels high; a 512-tap one 513; the deepest accepted text needs about 3 MB of
// stack, threads get 8 MB by default).
constexpr int kMaxDepth = 1536;
constexpr int kAbsorber = -1;   // result_types: a wire whose type the feedback cycle leaves open (flowz.hpp:542)

enum class Op : uint8_t {
    Placeholder,  // _k                       (k >= 1)
    Delay,        // _k[_n]  /  _k[-n]        (k >= 1, n >= 1)
    Const,        // literal terminal         (dtype, value)
    Param,        // std::ref(float) terminal (param index) -- scalar or per-channel at run time
    Neg, Add, Sub, Mul, Div,  // leaf arithmetic, evaluated with C++ built-in operator semantics
    Seq,          // L |= R   (also spelled L >> R)
    Par,          // L | R    input-splitting parallel
    Chan,         // (L , R)  fan-out
    Fb,           // ~x       unary feedback (user facing; removed by canonicalisation)
    Bfb           // binary_feedback(L = promise, R = future)  (private node)
};

struct Expr;
using ExprP = std::shared_ptr<const Expr>;

struct Expr {
    Op op;
    int k = 0;            // placeholder index / param index
    int n = 0;            // delay
    Dtype dtype = Dtype::F32;  // Const only
    double value = 0;     // Const only (already rounded to dtype)
    double imag = 0;      // Const of a complex dtype only
    int depth = 1;        // height of the tree below (and including) this node; bounded by kMaxDepth
    std::vector<ExprP> ch;
};

ExprP placeholder(int k);
ExprP delay(int k, int n);
ExprP constant(Dtype dt, double v, double imag = 0);
ExprP param(int idx);
ExprP unary(Op op, ExprP a);
ExprP binary(Op op, ExprP a, ExprP b);

bool is_terminal(const Expr& e);
bool is_arith(const Expr& e);

// ---- text form -------------------------------------------------------------------------------
// parse():  C++ operator precedence (postfix [] > unary - ~ > * / > + - > >> > | > |= (right
// assoc) > ,).  Literals: 2 (int), 0.5 (double), 0.5f (float), hex floats, $k (parameter k),
// bfb(L, R) builds the private binary feedback node, front(n) builds make_front<n>(), cplx{re,im} /
// cplxd{re,im} are std::complex<float> / <double> terminals (type analysis only).  Two spellings the reference only
// plans (TODO.md:8-9, 51-52) are accepted: _k<-n> == _k[_n], and expr[_n] == expr |= _1[_n] (every output delayed).
ExprP parse(const std::string& text);
std::string to_string(const Expr& e);       // fully parenthesised, round-trips through parse()
bool same_structure(const Expr& a, const Expr& b);

// ---- static analysis -------------------------------------------------------------------------
int input_arity(const Expr& e);
int output_arity(const Expr& e);
std::vector<int> max_input_delays(const Expr& e);
std::vector<int> min_input_delays(const Expr& e);   // -1 = wire unused
int n_params(const Expr& e);                        // 1 + highest $k used, 0 if none

// ResultType (flowz.hpp:515-644): the C++ type of every output wire of `e` when input wire k has type in[k-1]
// ((int)Dtype, or kAbsorber).  is_tuple = false when the reference yields a bare scalar (a terminal, a delayed
// placeholder or plain arithmetic: test/tests.cpp:198-199, 215) instead of a std::tuple.
struct ResultTypes {
    std::vector<int> types;
    bool is_tuple = false;
};
ResultTypes result_types(const Expr& e, const std::vector<int>& in);

// ---- canonicalisation ------------------------------------------------------------------------
ExprP make_front(int n);
ExprP add_front_panel(ExprP e);
ExprP make_canonical(ExprP e);                      // replaces every ~x by bfb(promise, future)
// compile(): front panel + canonical form.  arity 0 is accepted as an extension (the reference has
// no make_front<0>, TODO.md:65): the expression is canonicalised without a front panel.
// keep_feedback: 0 = every ~x must split into binary_feedback(promise, future) as in the reference (else Error);
// 1 = a ~x that cannot be split stays a unary feedback over its canonical body; 2 = every ~x stays whole.  Kept
// feedbacks are lowered with forward references (zg_ir.cpp) -- beyond the reference, which cannot compile them.
ExprP canonical_with_front(ExprP e, int keep_feedback = 0);

}  // namespace zg
