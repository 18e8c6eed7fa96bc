// zignal-b200 :: expression layer implementation.  See zg_expr.hpp for the reference map.
#include "zg_expr.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <sstream>

namespace zg {

// ------------------------------------------------------------------------------------------------
// constructors
// ------------------------------------------------------------------------------------------------

static ExprP mk(Expr e) { return std::make_shared<const Expr>(std::move(e)); }

ExprP placeholder(int k) {
    if (k < 1) throw Error("placeholder index must be >= 1");
    Expr e; e.op = Op::Placeholder; e.k = k; return mk(e);
}
ExprP delay(int k, int n) {
    if (k < 1) throw Error("placeholder index must be >= 1");
    if (n < 1) throw Error("delay must be >= 1 (write _k for the undelayed wire)");
    Expr e; e.op = Op::Delay; e.k = k; e.n = n; return mk(e);
}
ExprP constant(Dtype dt, double v, double imag) {
    Expr e; e.op = Op::Const; e.dtype = dt;
    switch (dt) {
        case Dtype::I32: e.value = (double)(int32_t)v; break;
        case Dtype::F32: e.value = (double)(float)v; break;
        case Dtype::F64: e.value = v; break;
        case Dtype::C64: e.value = (double)(float)v; e.imag = (double)(float)imag; break;
        case Dtype::C128: e.value = v; e.imag = imag; break;
    }
    return mk(e);
}
ExprP param(int idx) {
    if (idx < 0) throw Error("parameter index must be >= 0");
    Expr e; e.op = Op::Param; e.k = idx; return mk(e);
}
static void set_depth(Expr& e) {
    for (auto& c : e.ch) e.depth = std::max(e.depth, c->depth + 1);
    if (e.depth > kMaxDepth)
        throw Error("expression is nested more than " + std::to_string(kMaxDepth) + " levels deep");
}
ExprP unary(Op op, ExprP a) {
    Expr e; e.op = op; e.ch = {std::move(a)}; set_depth(e); return mk(e);
}
ExprP binary(Op op, ExprP a, ExprP b) {
    Expr e; e.op = op; e.ch = {std::move(a), std::move(b)}; set_depth(e); return mk(e);
}

bool is_terminal(const Expr& e) {
    return e.op == Op::Placeholder || e.op == Op::Const || e.op == Op::Param;
}
bool is_arith(const Expr& e) {
    return e.op == Op::Neg || e.op == Op::Add || e.op == Op::Sub || e.op == Op::Mul || e.op == Op::Div;
}

// ------------------------------------------------------------------------------------------------
// printer
// ------------------------------------------------------------------------------------------------

static void print_const(std::ostringstream& os, const Expr& e) {
    char buf[64];
    switch (e.dtype) {
        case Dtype::I32: std::snprintf(buf, sizeof buf, "%d", (int)e.value); break;
        case Dtype::F32: std::snprintf(buf, sizeof buf, "%af", e.value); break;   // hex float: exact
        case Dtype::F64: std::snprintf(buf, sizeof buf, "%a", e.value); break;
        case Dtype::C64: case Dtype::C128:
            std::snprintf(buf, sizeof buf, "%s{%a,%a}", e.dtype == Dtype::C64 ? "cplx" : "cplxd", e.value, e.imag);
            os << buf;
            return;
    }
    // a negative literal is printed as a parenthesised literal so that it is re-read as a Const,
    // not as Neg(Const)
    if (buf[0] == '-') os << "(" << buf << ")"; else os << buf;
}

static void print(std::ostringstream& os, const Expr& e) {
    auto bin = [&](const char* sym) {
        os << "("; print(os, *e.ch[0]); os << sym; print(os, *e.ch[1]); os << ")";
    };
    switch (e.op) {
        case Op::Placeholder: os << "_" << e.k; break;
        case Op::Delay: os << "_" << e.k << "[_" << e.n << "]"; break;
        case Op::Const: print_const(os, e); break;
        case Op::Param: os << "$" << e.k; break;
        case Op::Neg: os << "(-"; print(os, *e.ch[0]); os << ")"; break;
        case Op::Add: bin(" + "); break;
        case Op::Sub: bin(" - "); break;
        case Op::Mul: bin("*"); break;
        case Op::Div: bin("/"); break;
        case Op::Seq: bin(" |= "); break;
        case Op::Par: bin(" | "); break;
        case Op::Chan: bin(" , "); break;
        case Op::Fb: os << "(~"; print(os, *e.ch[0]); os << ")"; break;
        case Op::Bfb: os << "bfb("; print(os, *e.ch[0]); os << " , "; print(os, *e.ch[1]); os << ")"; break;
    }
}

std::string to_string(const Expr& e) {
    std::ostringstream os; print(os, e); return os.str();
}

bool same_structure(const Expr& a, const Expr& b) {
    if (a.op != b.op || a.k != b.k || a.n != b.n || a.ch.size() != b.ch.size()) return false;
    if (a.op == Op::Const && (a.dtype != b.dtype || std::memcmp(&a.value, &b.value, sizeof(double)) != 0 ||
                              std::memcmp(&a.imag, &b.imag, sizeof(double)) != 0))
        return false;
    for (size_t i = 0; i < a.ch.size(); ++i)
        if (!same_structure(*a.ch[i], *b.ch[i])) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------
// parser (precedence climbing, C++ operator precedence)
// ------------------------------------------------------------------------------------------------

namespace {

struct Parser {
    const std::string& s;
    size_t p = 0;
    explicit Parser(const std::string& text) : s(text) {}

    [[noreturn]] void fail(const std::string& msg) const {
        // (the text is echoed around the column only: expressions can be tens of kilobytes long)
        const size_t b = p > 60 ? p - 60 : 0, n = std::min<size_t>(s.size() - b, 120);
        throw Error("flowz parse error at column " + std::to_string(p + 1) + ": " + msg + "  in \"" + (b ? "..." : "") +
                    s.substr(b, n) + (b + n < s.size() ? "..." : "") + "\"");
    }
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) ++p; }
    bool eat(const char* tok) {
        ws();
        size_t n = std::strlen(tok);
        if (s.compare(p, n, tok) == 0) { p += n; return true; }
        return false;
    }
    bool peek(const char* tok) {
        ws();
        return s.compare(p, std::strlen(tok), tok) == 0;
    }
    void expect(const char* tok) { if (!eat(tok)) fail(std::string("expected '") + tok + "'"); }

    // wire / delay / parameter indices: bounded, so that no text can make the analyses or a voice's state
    // allocation overflow (a delay line of 2^20 samples is 4 MB of state per voice)
    static constexpr long kMaxIndex = 1l << 20;
    int integer() {
        ws();
        size_t b = p;
        long v = 0;
        while (p < s.size() && std::isdigit((unsigned char)s[p])) {
            v = v * 10 + (s[p] - '0');
            if (v > kMaxIndex) fail("index larger than " + std::to_string(kMaxIndex));
            ++p;
        }
        if (b == p) fail("expected integer");
        return (int)v;
    }

    // The parser recurses once per open parenthesis / right-nested `|=` (a chain of nine frames) and once per prefix
    // operator (one frame).  A tree level prints as at most one parenthesis and one prefix operator (to_string), so with
    // each kind bounded by the tree limit every printed tree can be read again.
    int nest[3] = {0, 0, 0};          // parentheses, prefix operators, right-nested |=
    struct Nest {
        Parser& ps;
        int kind;
        Nest(Parser& q, int k) : ps(q), kind(k) {
            if (++ps.nest[kind] > kMaxDepth + 8) ps.fail("nested more than " + std::to_string(kMaxDepth + 8) + " levels deep");
        }
        ~Nest() { --ps.nest[kind]; }
    };

    ExprP comma() {
        Nest guard(*this, 0);
        ExprP l = assign();
        while (peek(",")) { eat(","); l = binary(Op::Chan, l, assign()); }
        return l;
    }
    ExprP assign() {          // |= is right associative
        ExprP l = bitor_();
        if (eat("|=")) {
            Nest guard(*this, 2);
            return binary(Op::Seq, l, assign());
        }
        return l;
    }
    ExprP bitor_() {
        ExprP l = shift();
        for (;;) {
            ws();
            if (p < s.size() && s[p] == '|' && !(p + 1 < s.size() && s[p + 1] == '=')) {
                ++p; l = binary(Op::Par, l, shift());
            } else break;
        }
        return l;
    }
    ExprP shift() {           // >> : series spelling of the early prototypes, left associative
        ExprP l = additive();
        while (eat(">>")) l = binary(Op::Seq, l, additive());
        return l;
    }
    ExprP additive() {
        ExprP l = multiplicative();
        for (;;) {
            if (eat("+")) l = binary(Op::Add, l, multiplicative());
            else if (peek("-")) { eat("-"); l = binary(Op::Sub, l, multiplicative()); }
            else break;
        }
        return l;
    }
    ExprP multiplicative() {
        ExprP l = unary_();
        for (;;) {
            if (eat("*")) l = binary(Op::Mul, l, unary_());
            else if (eat("/")) l = binary(Op::Div, l, unary_());
            else break;
        }
        return l;
    }
    ExprP unary_() {
        if (!peek("~") && !peek("+") && !peek("-")) return postfix();
        Nest guard(*this, 1);                               // one level per prefix operator
        if (eat("~")) return unary(Op::Fb, unary_());
        if (eat("+")) return unary_();
        if (peek("-")) {
            eat("-");
            ws();
            // a minus sign directly in front of a literal is part of the literal, as in C++ where
            // `-0.3f * _1` multiplies the wire by the float constant -0.3f
            if (p < s.size() && (std::isdigit((unsigned char)s[p]) || s[p] == '.')) {
                ExprP num = number();
                ExprP c = constant(num->dtype, -num->value);
                return postfix_on(c);
            }
            return unary(Op::Neg, unary_());
        }
        return postfix();
    }
    ExprP postfix() { return postfix_on(primary()); }
    ExprP postfix_on(ExprP e) {
        ws();
        for (;;) {
            int n;
            if (p < s.size() && s[p] == '[') {          // _k[_n]  _k[-n]  _k[n]
                ++p;
                if (eat("_")) n = integer();
                else if (eat("-")) n = integer();
                else n = integer();
                expect("]");
            } else if (p + 1 < s.size() && s[p] == '<' && e->op == Op::Placeholder) {   // _k<-n>  (TODO.md:8-9)
                ++p;
                expect("-");
                n = integer();
                expect(">");
            } else break;
            if (e->op == Op::Placeholder) e = delay(e->k, n);
            else e = delay_expr(e, n);
            ws();
        }
        return e;
    }
    // expr[_n]: "(_1+_2)[_1] is equivalent to _1+_2 |= _1[_1]" (TODO.md:51-52); with m outputs every wire is delayed:
    // expr |= (_1[_n] | ... | _1[_n])
    ExprP delay_expr(const ExprP& e, int n) {
        const int m = output_arity(*e);
        if (m < 1) fail("cannot delay an expression without outputs");
        ExprP d = delay(1, n);
        for (int i = 1; i < m; ++i) d = binary(Op::Par, d, delay(1, n));
        return binary(Op::Seq, e, d);
    }
    ExprP number() {
        ws();
        const char* b = s.c_str() + p;
        char* end = nullptr;
        bool hex = (b[0] == '0' && (b[1] == 'x' || b[1] == 'X'));
        // decide int vs floating
        size_t q = p;
        bool floating = false;
        if (hex) {
            q += 2;
            while (q < s.size() && (std::isxdigit((unsigned char)s[q]) || s[q] == '.')) { if (s[q] == '.') floating = true; ++q; }
            if (q < s.size() && (s[q] == 'p' || s[q] == 'P')) floating = true;
        } else {
            while (q < s.size() && (std::isdigit((unsigned char)s[q]) || s[q] == '.')) { if (s[q] == '.') floating = true; ++q; }
            if (q < s.size() && (s[q] == 'e' || s[q] == 'E')) floating = true;
        }
        if (!floating && !(q < s.size() && (s[q] == 'f' || s[q] == 'F') && !hex)) {
            errno = 0;
            long long v = std::strtoll(b, &end, 0);
            if (errno == ERANGE || v > 2147483647ll || v < -2147483648ll) fail("integer literal does not fit a C++ int");
            p += (size_t)(end - b);
            return constant(Dtype::I32, (double)v);
        }
        double v = std::strtod(b, &end);
        if (end == b) fail("bad number");
        p += (size_t)(end - b);
        if (p < s.size() && (s[p] == 'f' || s[p] == 'F')) { ++p; return constant(Dtype::F32, (double)(float)v); }
        return constant(Dtype::F64, v);
    }
    ExprP primary() {
        ws();
        if (p >= s.size()) fail("unexpected end of expression");
        char c = s[p];
        if (c == '(') {
            ++p;
            ws();
            // "(-0x1.8p-1f)" : negative literal printed by to_string()
            if (p < s.size() && s[p] == '-') {
                size_t save = p;
                ++p; ws();
                if (p < s.size() && (std::isdigit((unsigned char)s[p]) || s[p] == '.')) {
                    ExprP num = number();
                    if (eat(")")) return constant(num->dtype, -num->value);
                }
                p = save;
            }
            ExprP e = comma();
            expect(")");
            return e;
        }
        if (c == '_') { ++p; return placeholder(integer()); }
        if (c == '$') { ++p; return param(integer()); }
        if (std::isdigit((unsigned char)c) || c == '.') return number();
        if (s.compare(p, 4, "bfb(") == 0) {
            p += 4;
            ExprP l = assign(); expect(","); ExprP r = assign(); expect(")");
            return binary(Op::Bfb, l, r);
        }
        if (s.compare(p, 5, "cplx{") == 0 || s.compare(p, 6, "cplxd{") == 0) {      // std::complex<float> / <double>
            const bool dbl = s[p + 4] == 'd';
            p += dbl ? 6 : 5;
            auto part = [&] {
                ws();
                const bool neg = p < s.size() && s[p] == '-';
                if (neg) ++p;
                ExprP n = number();
                return neg ? -n->value : n->value;
            };
            const double re = part();
            double im = 0;
            if (eat(",")) im = part();
            expect("}");
            return constant(dbl ? Dtype::C128 : Dtype::C64, re, im);
        }
        if (s.compare(p, 6, "front(") == 0) {
            p += 6; int n = integer(); expect(")");
            return make_front(n);
        }
        fail(std::string("unexpected character '") + c + "'");
    }
};

}  // namespace

ExprP parse(const std::string& text) {
    Parser ps(text);
    ExprP e = ps.comma();
    ps.ws();
    if (ps.p != text.size()) ps.fail("trailing characters");
    return e;
}

// ------------------------------------------------------------------------------------------------
// arity  (flowz/flowz.hpp:162-246)
// ------------------------------------------------------------------------------------------------

int input_arity(const Expr& e) {
    switch (e.op) {
        case Op::Delay:
        case Op::Placeholder: return e.k;
        case Op::Const:
        case Op::Param: return 0;
        case Op::Fb: return std::max(0, input_arity(*e.ch[0]) - output_arity(*e.ch[0]));
        case Op::Bfb:
            return std::max(0, input_arity(*e.ch[0]) - output_arity(*e.ch[1])) +
                   std::max(0, input_arity(*e.ch[1]) - output_arity(*e.ch[0]));
        case Op::Par: return input_arity(*e.ch[0]) + input_arity(*e.ch[1]);
        case Op::Seq:
            return input_arity(*e.ch[0]) + std::max(0, input_arity(*e.ch[1]) - output_arity(*e.ch[0]));
        default: {  // any other n-ary node (arithmetic, channel): max over the children
            int m = 0;
            for (auto& c : e.ch) m = std::max(m, input_arity(*c));
            return m;
        }
    }
}

int output_arity(const Expr& e) {
    switch (e.op) {
        case Op::Chan:
        case Op::Par: return output_arity(*e.ch[0]) + output_arity(*e.ch[1]);
        case Op::Fb: return output_arity(*e.ch[0]);
        case Op::Bfb: return output_arity(*e.ch[1]);
        case Op::Seq:
            return output_arity(*e.ch[1]) + std::max(0, output_arity(*e.ch[0]) - input_arity(*e.ch[1]));
        default: return 1;
    }
}

int n_params(const Expr& e) {
    int m = e.op == Op::Param ? e.k + 1 : 0;
    for (auto& c : e.ch) m = std::max(m, n_params(*c));
    return m;
}

// ------------------------------------------------------------------------------------------------
// per-wire delays  (flowz/flowz.hpp:286-506)
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// ResultType  (flowz/flowz.hpp:515-644)
//
// The reference computes, at C++ compile time, the type of every output wire from the types of the input wires.
// Inside a feedback the fed-back wires have no type yet, so it evaluates the body with a placeholder type
// `absorber` on those wires and lets every binary operation with an absorber operand take the type of the other
// operand (:536-548, applied in a second pass over the expression the first pass built, :572-589).  Here types are
// run-time values, so one pass with the rule  absorber op T = T op absorber = T  gives the same answer;
// absorber op absorber stays absorber (the reference's "leftover", TODO at :575-578: ~(_1[_1]) has no type).
// Unary minus keeps the type ("unary-op absorber -> absorber", :533).  Leaf arithmetic on actual types follows the
// usual arithmetic conversions of C++ (proto::_default, :640-642); std::complex<T> combines with T and with
// itself only, as operator* etc. of <complex> do.
// ------------------------------------------------------------------------------------------------

namespace {

const char* type_name(int t) {
    switch (t) {
        case (int)Dtype::I32: return "int";
        case (int)Dtype::F32: return "float";
        case (int)Dtype::F64: return "double";
        case (int)Dtype::C64: return "std::complex<float>";
        case (int)Dtype::C128: return "std::complex<double>";
        default: return "absorber";
    }
}

int arith_type(int a, int b) {
    if (a == kAbsorber) return b;                                  // absorb_left  :539-542
    if (b == kAbsorber) return a;                                  // absorb_right :544-545
    auto cplx = [](int t) { return t == (int)Dtype::C64 || t == (int)Dtype::C128; };
    auto base = [](int t) { return t == (int)Dtype::C64 ? (int)Dtype::F32 : (int)Dtype::F64; };
    if (cplx(a) || cplx(b)) {
        const bool ok = cplx(a) && cplx(b) ? a == b : cplx(a) ? b == base(a) : a == base(b);
        if (!ok)
            throw Error(std::string("no arithmetic operator for ") + type_name(a) + " and " + type_name(b) +
                        " (std::complex<T> combines with T and std::complex<T> only)");
        return cplx(a) ? a : b;
    }
    return std::max(a, b);                                         // int < float < double
}

std::vector<int> take_types(const std::vector<int>& v, int n) {   // tuple_take, tuple_tools.hpp:92-103
    return n >= (int)v.size() ? v : std::vector<int>(v.begin(), v.begin() + std::max(n, 0));
}
std::vector<int> drop_types(const std::vector<int>& v, int n) {   // tuple_drop, tuple_tools.hpp:138-150
    return n >= (int)v.size() ? std::vector<int>{} : std::vector<int>(v.begin() + std::max(n, 0), v.end());
}

ResultTypes rt(const Expr& e, const std::vector<int>& st) {
    auto wire = [&](int k) {                                       // get_fn :550-557
        if (k > (int)st.size())
            throw Error("ResultType: _" + std::to_string(k) + " reads past the " + std::to_string(st.size()) +
                        " typed wires available at this point");
        return ResultTypes{{st[k - 1]}, false};
    };
    auto with_absorbers = [&](int n) {                             // repeat_fn + tuple_cat_fn :600, :607
        std::vector<int> s2((size_t)n, kAbsorber);
        s2.insert(s2.end(), st.begin(), st.end());
        return s2;
    };
    auto tuple = [](ResultTypes r) { r.is_tuple = true; return r; };   // make_flat_tuple :559-566
    switch (e.op) {
        case Op::Delay:                                            // :582-585
        case Op::Placeholder: return wire(e.k);                    // :586-589
        case Op::Const: return {{(int)e.dtype}, false};            // :590-593
        case Op::Param: return {{(int)Dtype::F32}, false};         // std::ref(float): converts to float in arithmetic
        case Op::Bfb: return tuple(rt(*e.ch[1], with_absorbers(output_arity(*e.ch[0]))));   // :594-609
        case Op::Fb: return tuple(rt(*e.ch[0], with_absorbers(output_arity(*e.ch[0]))));    // :610-615
        case Op::Seq: {                                            // :616-624
            const std::vector<int> l = rt(*e.ch[0], st).types;
            const int n = input_arity(*e.ch[1]);
            ResultTypes r = tuple(rt(*e.ch[1], take_types(l, n)));
            const std::vector<int> around = drop_types(l, n);
            r.types.insert(r.types.end(), around.begin(), around.end());
            return r;
        }
        case Op::Par: {                                            // :625-630
            const int n = input_arity(*e.ch[0]);
            ResultTypes r = tuple(rt(*e.ch[0], take_types(st, n)));
            const std::vector<int> b = rt(*e.ch[1], drop_types(st, n)).types;
            r.types.insert(r.types.end(), b.begin(), b.end());
            return r;
        }
        case Op::Chan: {                                           // :631-636
            ResultTypes r = tuple(rt(*e.ch[0], st));
            const std::vector<int> b = rt(*e.ch[1], st).types;
            r.types.insert(r.types.end(), b.begin(), b.end());
            return r;
        }
        default: {                                                 // leaf arithmetic, proto::_default :637-641
            int t = kAbsorber;
            for (size_t i = 0; i < e.ch.size(); ++i) {
                const ResultTypes c = rt(*e.ch[i], st);
                if (c.is_tuple) throw Error("ResultType: arithmetic on a tuple of wires (operands must be single wires)");
                t = i == 0 ? c.types[0] : arith_type(t, c.types[0]);
            }
            return {{t}, false};
        }
    }
}

}  // namespace

ResultTypes result_types(const Expr& e, const std::vector<int>& in) { return rt(e, in); }

namespace {

std::vector<int> drop(const std::vector<int>& v, int n) {       // tuple_drop: N > size -> empty
    if (n >= (int)v.size()) return {};
    return std::vector<int>(v.begin() + n, v.end());
}
std::vector<int> cat(std::vector<int> a, const std::vector<int>& b) {
    a.insert(a.end(), b.begin(), b.end());
    return a;
}
int map_min(int n, int m) { return n == -1 ? m : m == -1 ? n : std::min(n, m); }

// zip over the common prefix, then the leftover of whichever side is longer (:364-379, :403-418)
template <class F>
std::vector<int> zip_wires(const std::vector<int>& a, const std::vector<int>& b, F f) {
    size_t m = std::min(a.size(), b.size());
    std::vector<int> r;
    for (size_t i = 0; i < m; ++i) r.push_back(f(a[i], b[i]));
    for (size_t i = m; i < a.size(); ++i) r.push_back(a[i]);
    for (size_t i = m; i < b.size(); ++i) r.push_back(b[i]);
    return r;
}

template <bool kMin>
std::vector<int> input_delays(const Expr& e) {
    const int other = kMin ? -1 : 0;
    switch (e.op) {
        case Op::Delay: {
            std::vector<int> r(e.k, other); r[e.k - 1] = e.n; return r;
        }
        case Op::Placeholder: {
            std::vector<int> r(e.k, other); r[e.k - 1] = 0; return r;
        }
        case Op::Const:
        case Op::Param: return {};
        case Op::Fb: return drop(input_delays<kMin>(*e.ch[0]), output_arity(*e.ch[0]));
        case Op::Bfb:
            return cat(drop(input_delays<kMin>(*e.ch[0]), output_arity(*e.ch[1])),
                       drop(input_delays<kMin>(*e.ch[1]), output_arity(*e.ch[0])));
        case Op::Par: return cat(input_delays<kMin>(*e.ch[0]), input_delays<kMin>(*e.ch[1]));
        case Op::Seq:
            return cat(input_delays<kMin>(*e.ch[0]),
                       drop(input_delays<kMin>(*e.ch[1]), output_arity(*e.ch[0])));
        default: {
            std::vector<int> acc;
            for (auto& c : e.ch) {
                auto d = input_delays<kMin>(*c);
                acc = kMin ? zip_wires(d, acc, map_min)
                           : zip_wires(d, acc, [](int a, int b) { return std::max(a, b); });
            }
            return acc;
        }
    }
}

}  // namespace

std::vector<int> max_input_delays(const Expr& e) { return input_delays<false>(e); }
std::vector<int> min_input_delays(const Expr& e) { return input_delays<true>(e); }

// ------------------------------------------------------------------------------------------------
// front panel + canonical form  (flowz/flowz.hpp:261-277, 794-935)
// ------------------------------------------------------------------------------------------------

ExprP make_front(int n) {
    if (n < 1) throw Error("make_front<0> does not exist (flowz.hpp:261-271)");
    ExprP f = placeholder(1);
    for (int i = 1; i < n; ++i) f = binary(Op::Par, f, placeholder(1));
    return f;
}

ExprP add_front_panel(ExprP e) {
    int n = input_arity(*e);
    return binary(Op::Seq, make_front(n), e);
}

namespace {

// the predicate used by split_future_subexpr (:869-874): does r need any of its first n input
// wires undelayed?
bool needs_num_direct_input(const Expr& r, int n) {
    auto d = min_input_delays(r);
    int m = std::min<int>(n, (int)d.size());          // tuple_take<N>: N > size -> whole tuple
    for (int i = 0; i < m; ++i) if (d[i] == 0) return true;
    return false;
}

using Split = std::vector<ExprP>;   // 1 element = unsplit, 2 elements = (promise side, future side)

Split u2b(const ExprP& e);

Split split_in_sequence(const ExprP& l, const ExprP& r) {     // :887-935
    Split ul = u2b(l);
    if (ul.size() == 2) return {ul[0], binary(Op::Seq, ul[1], r)};
    const ExprP& l1 = ul[0];
    if (needs_num_direct_input(*r, output_arity(*l1))) {
        Split ur = u2b(r);
        if (ur.size() == 2) return {binary(Op::Seq, l1, ur[0]), ur[1]};
        return {binary(Op::Seq, l1, ur[0])};
    }
    return {l1, r};
}

Split u2b(const ExprP& e) {                                   // :811-846
    if (is_terminal(*e)) return {e};
    if (e->op == Op::Seq) return split_in_sequence(e->ch[0], e->ch[1]);
    // lifted_default: rebuild the node over the first elements, keep the tail of the first child
    std::vector<Split> res;
    for (auto& c : e->ch) res.push_back(u2b(c));
    Expr n = *e;
    n.ch.clear();
    for (auto& r : res) n.ch.push_back(r[0]);
    Split out{mk(n)};
    if (!res.empty()) for (size_t i = 1; i < res[0].size(); ++i) out.push_back(res[0][i]);
    return out;
}

// compile() only (canonical_with_front): for a graph the reference cannot compile because of a feedback -- nested
// feedback whose inner loop takes the outer fed-back wire in its promise part, parallel combiners inside a loop,
// TODO.md:11-29 and the disabled test/tests.cpp:59 -- every `~x` stays a unary feedback over the canonical x (mode 2;
// mode 1 keeps only those that do not split).  The lowering resolves the fed-back wires of such a node as forward
// references and refuses only genuine zero-delay loops (zg_ir.cpp).  make_canonical() itself stays strict.
thread_local int g_keep_feedback = 0;

ExprP split_future_subexpr(const ExprP& x) {                  // :862-884
    if (g_keep_feedback == 2) return unary(Op::Fb, make_canonical(x));
    ExprP chain = binary(Op::Seq, make_front(output_arity(*x)), x);
    Split s = u2b(make_canonical(chain));
    if (s.size() != 2 && g_keep_feedback == 1) return unary(Op::Fb, make_canonical(x));
    if (s.size() != 2)
        throw Error("feedback ~(" + to_string(*x) +
                    ") cannot be split into a promise and a future part: every path from the "
                    "fed-back wires needs a direct (undelayed) input (flowz.hpp:879-883)");
    return binary(Op::Bfb, s[0], s[1]);
}

}  // namespace

ExprP make_canonical(ExprP e) {                               // :794-805
    if (is_terminal(*e)) return e;
    if (e->op == Op::Fb) return split_future_subexpr(e->ch[0]);
    Expr n = *e;
    for (auto& c : n.ch) c = make_canonical(c);
    return mk(n);
}

ExprP canonical_with_front(ExprP e, int keep_feedback) {      // compile(), :1233-1249
    struct Mode {
        int old = g_keep_feedback;
        explicit Mode(int m) { g_keep_feedback = m; }
        ~Mode() { g_keep_feedback = old; }
    } mode(keep_feedback);
    if (input_arity(*e) == 0) return make_canonical(e);       // extension, see header
    return make_canonical(add_front_panel(e));
}

}  // namespace zg
