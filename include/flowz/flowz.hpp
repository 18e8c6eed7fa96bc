// flowz -- C++ EDSL surface of zignal-b200 (header-only, C++17, no Boost).
//
// Source-level stand-in for the reference's <flowz/flowz.hpp> (andre-bergner/zignal,
// /root/reference/flowz/flowz.hpp): the same spellings build the same graphs --
//
//     _1 ... _6, make_placeholder<n>()        :75-82, 1252-1257
//     make_terminal(x), std::ref(param)       :68-72, flowz/README.md:42-63
//     a |= b  (and a >> b)   series           :92   (>>: experimental_steps/wires_mono_only.cpp:37)
//     a | b                  parallel         :91
//     (a , b)                fan-out          :90
//     ~a                     feedback         :93
//     _k[_n]  (and _k[-n])   unit delays      :84-85 (_[-n]: experimental_steps/delay_expression.cpp:99)
//     + - * / unary -        leaf arithmetic  :769-772
//     compile(expr)                           :1233-1249
//     f(x1, ..., xN) -> std::tuple<...>, currying with fewer arguments   :1193-1229
//     input_arity, output_arity, max_input_delays, transforms::make_canonical,
//     make_binary_feedback                    :162-246, 443-506, 794-805, 100-102
//
// -- but nothing is evaluated by template expansion.  An expression is a small typed tree whose
// only compile-time content is its wire counts (needed for the std::tuple return type and for the
// argument-count check).  compile() sends the tree as text through the C ABI
// (include/zignal_b200.h); the library canonicalises it, lowers it to a flat tick program and
//   * ticks it on the host for the scalar operator() (one voice, one sample -- the reference's
//     whole API), and
//   * evaluates it block-wise for many channels on a B200 through block_evaluator (new).
//
// The tuple operator() returns has the reference's per-wire C++ types: every node carries a constexpr rule
// (`types`) that propagates the argument types the way the reference's evaluators do -- C++'s usual arithmetic
// conversions on the leaves, delayed reads are float because the state is (flowz.hpp:136, :1245), fed-back wires
// take the type of what is fed back (a fixed point, the role of the reference's `absorber`, :523-548) -- and the
// tick checks it against the types the library computed at run time.
// Known deviation from the reference (documented in DESIGN.md): canonical forms are compared with operator== on
// run-time trees, not with std::is_same on types.
#pragma once

#include <array>
#include <cstdint>
#include <complex>
#include <cstdio>
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../zignal_b200.h"

namespace flowz {

struct error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace detail {

inline void check(int status) {
    if (status != ZG_OK) throw error(zg_last_error());
}

constexpr int cmax(int a, int b) { return a > b ? a : b; }

// ---- compile-time result types of one tick ----------------------------------------------------
// Type codes are zg_dtype (ZG_I32 < ZG_F32 < ZG_F64: promotion is the maximum); kOpen = not typed yet (a fed-back
// wire before the fixed point has reached it), which any other type absorbs.
constexpr int kOpen = -1;
constexpr int promote(int a, int b) { return a == kOpen ? b : b == kOpen ? a : cmax(a, b); }
template <size_t N> using type_list = std::array<int, N>;
template <size_t A, size_t B>
constexpr type_list<A + B> cat(const type_list<A>& a, const type_list<B>& b) {
    type_list<A + B> r{};
    for (size_t i = 0; i < A; ++i) r[i] = a[i];
    for (size_t i = 0; i < B; ++i) r[A + i] = b[i];
    return r;
}
template <size_t K, size_t N>
constexpr type_list<(K < N ? K : N)> take(const type_list<N>& a) {       // tuple_tools.hpp:92-103
    type_list<(K < N ? K : N)> r{};
    for (size_t i = 0; i < r.size(); ++i) r[i] = a[i];
    return r;
}
template <size_t K, size_t N>
constexpr type_list<(K < N ? N - K : 0)> drop(const type_list<N>& a) {   // tuple_tools.hpp:138-150
    type_list<(K < N ? N - K : 0)> r{};
    for (size_t i = 0; i < r.size(); ++i) r[i] = a[K + i];
    return r;
}
template <size_t N>
constexpr type_list<N> filled(int v) {
    type_list<N> r{};
    for (size_t i = 0; i < N; ++i) r[i] = v;
    return r;
}
template <class T> constexpr int type_code() {
    return std::is_integral<T>::value ? ZG_I32 : std::is_same<T, float>::value ? ZG_F32 : ZG_F64;
}
template <int Code> struct code_type { using type = float; };
template <> struct code_type<ZG_I32> { using type = int; };
template <> struct code_type<ZG_F64> { using type = double; };

// writer state: parameters ($k) are numbered in order of first appearance
struct writer {
    std::ostringstream os;
    std::vector<const float*> refs;
    int ref_index(const float* p) {
        for (size_t i = 0; i < refs.size(); ++i)
            if (refs[i] == p) return (int)i;
        refs.push_back(p);
        return (int)refs.size() - 1;
    }
};

struct expr_tag {};
template <class T> constexpr bool is_expr_v = std::is_base_of<expr_tag, std::decay_t<T>>::value;

namespace tag {
struct plus {}; struct minus {}; struct multiplies {}; struct divides {}; struct negate {};
struct sequence {}; struct parallel {}; struct channel {}; struct feedback {}; struct binary_feedback {};
}

template <class Tag> struct symbol;
template <> struct symbol<tag::plus> { static constexpr const char* s = " + "; };
template <> struct symbol<tag::minus> { static constexpr const char* s = " - "; };
template <> struct symbol<tag::multiplies> { static constexpr const char* s = "*"; };
template <> struct symbol<tag::divides> { static constexpr const char* s = "/"; };
template <> struct symbol<tag::sequence> { static constexpr const char* s = " |= "; };
template <> struct symbol<tag::parallel> { static constexpr const char* s = " | "; };
template <> struct symbol<tag::channel> { static constexpr const char* s = " , "; };

// ---- terminals --------------------------------------------------------------------------------

template <int N> struct placeholder_expr;

template <int K>
struct delayed_expr : expr_tag {
    static constexpr int in = K, out = 1;
    static constexpr bool has_double = false;
    int n;
    void write(writer& w) const { w.os << "_" << K << "[_" << n << "]"; }
    template <size_t N> static constexpr type_list<1> types(const type_list<N>&) { return {ZG_F32}; }   // the state is float
};

template <int K>
struct placeholder_expr : expr_tag {
    static_assert(K >= 1, "placeholders are numbered from 1");
    static constexpr int in = K, out = 1;
    static constexpr bool has_double = false;
    void write(writer& w) const { w.os << "_" << K; }
    template <size_t N> static constexpr type_list<1> types(const type_list<N>& t) {
        static_assert((size_t)K <= N, "placeholder reads past the wires it is given");
        return {t[K - 1]};
    }
    template <int N>
    delayed_expr<K> operator[](placeholder_expr<N>) const { return {{}, N}; }    // _k[_n]
    delayed_expr<K> operator[](int n) const {                                    // _k[-n]
        if (n == 0) throw error("a delay of 0 is the wire itself");
        return {{}, n < 0 ? -n : n};
    }
};

template <class T> struct is_complex : std::false_type {};
template <> struct is_complex<std::complex<float>> : std::true_type {};
template <> struct is_complex<std::complex<double>> : std::true_type {};

template <class T>
struct literal_expr : expr_tag {
    static constexpr int in = 0, out = 1;
    static constexpr bool has_double = std::is_same<T, double>::value;
    T value;
    template <size_t N> static constexpr type_list<1> types(const type_list<N>&) {
        static_assert(!is_complex<T>::value, "complex terminals are typed by transforms::ResultType, not evaluated (flowz.hpp:1245)");
        return {type_code<T>()};
    }
    void write(writer& w) const {
        char buf[64];
        if constexpr (is_complex<T>::value) {          // typed by ResultType, not evaluated (flowz.hpp:1245)
            std::snprintf(buf, sizeof buf, "%s{%a,%a}", sizeof(T) == 8 ? "cplx" : "cplxd", (double)value.real(), (double)value.imag());
            w.os << buf;
            return;
        } else if constexpr (std::is_integral<T>::value) std::snprintf(buf, sizeof buf, "%d", (int)value);
        else if constexpr (std::is_same<T, float>::value) std::snprintf(buf, sizeof buf, "%af", (double)value);
        else std::snprintf(buf, sizeof buf, "%a", (double)value);
        if (buf[0] == '-') w.os << "(" << buf << ")"; else w.os << buf;
    }
};

struct ref_expr : expr_tag {                       // std::ref(x): non-owning, read at every tick
    static constexpr int in = 0, out = 1;
    static constexpr bool has_double = false;
    const float* p;
    void write(writer& w) const { w.os << "$" << w.ref_index(p); }
    template <size_t N> static constexpr type_list<1> types(const type_list<N>&) { return {ZG_F32}; }
};

// ---- operator nodes ---------------------------------------------------------------------------

template <class Tag, class A, class B> struct binary_expr;

// expr[_n] for a one-output expression: "(_1+_2)[_1] is equivalent to _1+_2 |= _1[_1]" (reference TODO.md:51-52)
template <class Self>
struct delayable {
    template <int N>
    auto operator[](placeholder_expr<N>) const { return delayed_by(N); }
    auto operator[](int n) const {
        if (n == 0) throw error("a delay of 0 is the wire itself");
        return delayed_by(n < 0 ? -n : n);
    }
private:
    auto delayed_by(int n) const {
        static_assert(Self::out == 1, "expr[_n] needs an expression with one output");
        return binary_expr<tag::sequence, Self, delayed_expr<1>>{{}, {}, static_cast<const Self&>(*this), delayed_expr<1>{{}, n}};
    }
};

template <class Tag, class A>
struct unary_expr : expr_tag, delayable<unary_expr<Tag, A>> {
    A a;
    static constexpr int in = std::is_same<Tag, tag::feedback>::value ? cmax(0, A::in - A::out) : A::in;
    static constexpr int out = std::is_same<Tag, tag::feedback>::value ? A::out : 1;
    static constexpr bool has_double = A::has_double;
    void write(writer& w) const {
        w.os << (std::is_same<Tag, tag::feedback>::value ? "(~" : "(-");
        a.write(w);
        w.os << ")";
    }
    // ~a: the first out(a) inputs of a are its own outputs.  Their types are the least fixed point of a's own rule,
    // started from "open" (three rounds reach it: int < float < double); a wire nothing ever types reads as float.
    template <size_t N> static constexpr type_list<(size_t)out> types(const type_list<N>& t) {
        if constexpr (std::is_same<Tag, tag::feedback>::value) {
            type_list<(size_t)A::out> fed = filled<(size_t)A::out>(kOpen);
            for (int round = 0; round < 4; ++round) fed = A::types(cat(fed, t));
            for (size_t i = 0; i < fed.size(); ++i) if (fed[i] == kOpen) fed[i] = ZG_F32;
            return fed;
        } else {
            return A::types(t);
        }
    }
};

template <class Tag, class A, class B>
struct binary_expr : expr_tag, delayable<binary_expr<Tag, A, B>> {
    A a;
    B b;
    static constexpr bool has_double = A::has_double || B::has_double;
    static constexpr int in =
        std::is_same<Tag, tag::sequence>::value ? A::in + cmax(0, B::in - A::out)
        : std::is_same<Tag, tag::parallel>::value ? A::in + B::in
        : std::is_same<Tag, tag::binary_feedback>::value ? cmax(0, A::in - B::out) + cmax(0, B::in - A::out)
        : cmax(A::in, B::in);
    static constexpr int out =
        std::is_same<Tag, tag::sequence>::value ? B::out + cmax(0, A::out - B::in)
        : (std::is_same<Tag, tag::parallel>::value || std::is_same<Tag, tag::channel>::value) ? A::out + B::out
        : std::is_same<Tag, tag::binary_feedback>::value ? B::out
        : 1;
    void write(writer& w) const {
        if constexpr (std::is_same<Tag, tag::binary_feedback>::value) {
            w.os << "bfb("; a.write(w); w.os << " , "; b.write(w); w.os << ")";
        } else {
            w.os << "("; a.write(w); w.os << symbol<Tag>::s; b.write(w); w.os << ")";
        }
    }
    // the routing rules of the reference's evaluators (sequence :960-1001, parallel :1076-1101, channel :765-768,
    // binary_feedback :1031-1074), on types instead of values
    template <size_t N> static constexpr type_list<(size_t)out> types(const type_list<N>& t) {
        if constexpr (std::is_same<Tag, tag::sequence>::value) {
            const auto l = A::types(t);
            const auto r = B::types(cat(l, drop<(size_t)A::in>(t)));
            return cat(r, drop<(size_t)B::in>(l));
        } else if constexpr (std::is_same<Tag, tag::parallel>::value) {
            return cat(A::types(take<(size_t)A::in>(t)), B::types(drop<(size_t)A::in>(t)));
        } else if constexpr (std::is_same<Tag, tag::channel>::value) {
            return cat(A::types(t), B::types(t));
        } else if constexpr (std::is_same<Tag, tag::binary_feedback>::value) {
            return B::types(cat(filled<(size_t)A::out>(ZG_F32), t));      // the future part reads the fed-back wires delayed
        } else {
            return {promote(A::types(t)[0], B::types(t)[0])};            // C++'s usual arithmetic conversions
        }
    }
};

// ---- lifting plain values into terminals ------------------------------------------------------

template <class T, class = void> struct as_expr_impl;
template <class T>
struct as_expr_impl<T, std::enable_if_t<is_expr_v<T>>> {
    using type = std::decay_t<T>;
    static type make(const T& t) { return t; }
};
template <class T>
struct as_expr_impl<T, std::enable_if_t<std::is_arithmetic<std::decay_t<T>>::value>> {
    using V = std::conditional_t<std::is_integral<std::decay_t<T>>::value, int,
              std::conditional_t<std::is_same<std::decay_t<T>, float>::value, float, double>>;
    using type = literal_expr<V>;
    static type make(const T& t) { return {{}, (V)t}; }
};
template <class T>
struct as_expr_impl<T, std::enable_if_t<is_complex<std::decay_t<T>>::value>> {
    using type = literal_expr<std::decay_t<T>>;
    static type make(const T& t) { return {{}, t}; }
};
template <class T>
struct as_expr_impl<std::reference_wrapper<T>, void> {
    static_assert(std::is_same<std::remove_const_t<T>, float>::value, "std::ref parameters must be float");
    using type = ref_expr;
    static type make(std::reference_wrapper<T> r) { return {{}, &r.get()}; }
};
template <class T> using as_expr_t = typename as_expr_impl<std::decay_t<T>>::type;
template <class T> as_expr_t<T> as_expr(const T& t) { return as_expr_impl<std::decay_t<T>>::make(t); }

template <class T> constexpr bool is_operand_v =
    is_expr_v<T> || std::is_arithmetic<std::decay_t<T>>::value || is_complex<std::decay_t<T>>::value;
template <class T> struct is_refw : std::false_type {};
template <class T> struct is_refw<std::reference_wrapper<T>> : std::true_type {};
template <class A, class B> constexpr bool arith_ok_v =
    (is_expr_v<A> && (is_operand_v<B> || is_refw<std::decay_t<B>>::value)) ||
    (is_expr_v<B> && (is_operand_v<A> || is_refw<std::decay_t<A>>::value));

template <class Tag, class A, class B>
binary_expr<Tag, as_expr_t<A>, as_expr_t<B>> make_binary(const A& a, const B& b) {
    return {{}, {}, as_expr(a), as_expr(b)};
}

template <class E>
std::string to_text(const E& e, std::vector<const float*>* refs = nullptr) {
    writer w;
    e.write(w);
    if (refs) *refs = w.refs;
    return w.os.str();
}

}  // namespace detail

// ---- operators (found by ADL on detail:: types; also visible via `using namespace flowz`) --------

namespace detail {

#define FLOWZ_ARITH(op, T)                                                           \
    template <class A, class B, class = std::enable_if_t<arith_ok_v<A, B>>>          \
    auto operator op(const A& a, const B& b) { return make_binary<tag::T>(a, b); }
FLOWZ_ARITH(+, plus)
FLOWZ_ARITH(-, minus)
FLOWZ_ARITH(*, multiplies)
FLOWZ_ARITH(/, divides)
#undef FLOWZ_ARITH

#define FLOWZ_COMB(op, T)                                                                      \
    template <class A, class B, class = std::enable_if_t<is_expr_v<A> && is_expr_v<B>>>       \
    auto operator op(const A& a, const B& b) { return make_binary<tag::T>(a, b); }
FLOWZ_COMB(|=, sequence)
FLOWZ_COMB(>>, sequence)
FLOWZ_COMB(|, parallel)
#undef FLOWZ_COMB

template <class A, class B, class = std::enable_if_t<is_expr_v<A> && is_expr_v<B>>>
auto operator,(const A& a, const B& b) { return make_binary<tag::channel>(a, b); }

template <class A, class = std::enable_if_t<is_expr_v<A>>>
unary_expr<tag::negate, A> operator-(const A& a) { return {{}, {}, a}; }
template <class A, class = std::enable_if_t<is_expr_v<A>>>
unary_expr<tag::feedback, A> operator~(const A& a) { return {{}, {}, a}; }

}  // namespace detail

// ---- building blocks ------------------------------------------------------------------------------

template <int n>
detail::placeholder_expr<n> make_placeholder() { return {}; }

template <class X>
auto make_terminal(X x) { return detail::as_expr(x); }

template <class L, class R>
auto make_binary_feedback(const L& l, const R& r) { return detail::make_binary<detail::tag::binary_feedback>(l, r); }

const auto _1 = make_placeholder<1>();
const auto _2 = make_placeholder<2>();
const auto _3 = make_placeholder<3>();
const auto _4 = make_placeholder<4>();
const auto _5 = make_placeholder<5>();
const auto _6 = make_placeholder<6>();

template <class E> std::string to_string(const E& e) { return detail::to_text(e); }

// ---- static analysis function objects (usable like the reference's transforms) -------------------

struct input_arity {
    template <class E> int operator()(const E& e) const {
        int n = 0; detail::check(zg_expr_arity(detail::to_text(e).c_str(), &n, nullptr)); return n;
    }
};
struct output_arity {
    template <class E> int operator()(const E& e) const {
        int n = 0; detail::check(zg_expr_arity(detail::to_text(e).c_str(), nullptr, &n)); return n;
    }
};
namespace detail {
template <class E> std::vector<int> delays_of(const E& e, int which) {
    int buf[64], n = 0;
    check(zg_expr_delays(to_text(e).c_str(), which, buf, 64, &n));
    return std::vector<int>(buf, buf + (n < 64 ? n : 64));
}
}
struct max_input_delays {
    template <class E> std::vector<int> operator()(const E& e) const { return detail::delays_of(e, 0); }
};
struct min_input_delays {
    template <class E> std::vector<int> operator()(const E& e) const { return detail::delays_of(e, 1); }
};

// A run-time expression (result of a rewrite).  Compares structurally with any expression.
struct dyn_expr : detail::expr_tag {
    std::string text;
    void write(detail::writer& w) const { w.os << text; }
};
template <class A, class B, class = std::enable_if_t<detail::is_expr_v<A> && detail::is_expr_v<B>>>
bool same_expr(const A& a, const B& b) {
    // normalise both sides through the library's parser/printer
    auto norm = [](const std::string& s) {
        std::vector<char> buf(s.size() * 4 + 256);
        // make_canonical is the identity on trees without '~'
        detail::check(zg_expr_canonical(s.c_str(), buf.data(), buf.size()));
        return std::string(buf.data());
    };
    return norm(detail::to_text(a)) == norm(detail::to_text(b));
}

namespace transforms {
using flowz::input_arity;
using flowz::output_arity;
using flowz::max_input_delays;
using flowz::min_input_delays;
// ResultType (flowz.hpp:515-644).  The reference returns a *value of the result type*, to be inspected with
// decltype; here the types are run-time data: r(expr, std::tuple<float>{}).is<std::tuple<float, double>>().
struct result_type_info {
    std::vector<int> types;                              // zg_dtype per output wire (ZG_TYPE_OPEN = leftover absorber)
    bool is_tuple = false;                               // false: the reference yields a bare scalar (test/tests.cpp:198)
    template <class T> bool is() const { return match(static_cast<T*>(nullptr)); }

private:
    template <class T> static constexpr int code() {
        return std::is_same<T, int>::value ? ZG_I32 : std::is_same<T, float>::value ? ZG_F32
             : std::is_same<T, double>::value ? ZG_F64 : std::is_same<T, std::complex<float>>::value ? ZG_C64
             : std::is_same<T, std::complex<double>>::value ? ZG_C128 : -2;
    }
    template <class T> bool match(T*) const { return !is_tuple && types.size() == 1 && types[0] == code<T>(); }
    template <class... Ts> bool match(std::tuple<Ts...>*) const {
        const int want[] = {code<Ts>()..., 0};
        if (!is_tuple || types.size() != sizeof...(Ts)) return false;
        for (size_t i = 0; i < sizeof...(Ts); ++i) if (types[i] != want[i]) return false;
        return true;
    }
};
struct ResultType {
    template <class E, class... Ts> result_type_info operator()(const E& e, const std::tuple<Ts...>&) const {
        const int in[] = {result_type_code<Ts>()..., 0};
        int out[64], n = 0, tup = 0;
        detail::check(zg_expr_result_types(detail::to_text(e).c_str(), in, (int)sizeof...(Ts), out, 64, &n, &tup));
        result_type_info r;
        r.types.assign(out, out + (n < 64 ? n : 64));
        r.is_tuple = tup != 0;
        return r;
    }

private:
    template <class T> static constexpr int result_type_code() {
        return std::is_same<T, int>::value ? ZG_I32 : std::is_same<T, float>::value ? ZG_F32
             : std::is_same<T, double>::value ? ZG_F64 : std::is_same<T, std::complex<float>>::value ? ZG_C64 : ZG_C128;
    }
};
struct make_canonical {                                  // flowz.hpp:794-805
    template <class E> dyn_expr operator()(const E& e) const {
        std::string s = detail::to_text(e);
        std::vector<char> buf(s.size() * 8 + 1024);
        detail::check(zg_expr_canonical(s.c_str(), buf.data(), buf.size()));
        dyn_expr d; d.text = buf.data(); return d;
    }
};
}  // namespace transforms

// ---- block evaluator: `channels` voices of one compiled graph on one B200 ----------------------------

class block_evaluator {
    std::shared_ptr<zg_plan> plan_;
public:
    block_evaluator() = default;
    explicit block_evaluator(zg_plan* p) : plan_(p, zg_plan_destroy) {}
    zg_plan* handle() const { return plan_.get(); }
    // device pointers, planar [channels][ld] (or interleaved, per the plan), asynchronous on `stream`
    void process(const float* const* in, float* const* out, int64_t n_samples, int64_t ld_in,
                 int64_t ld_out, void* stream = nullptr) {
        detail::check(zg_process(plan_.get(), (const void* const*)in, (void* const*)out, n_samples, ld_in, ld_out, stream));
    }
    // host pointers: H2D, kernel, D2H, synchronise
    void process_host(const float* const* in, float* const* out, int64_t n_samples, int64_t ld_in, int64_t ld_out) {
        detail::check(zg_process_host(plan_.get(), (const void* const*)in, (void* const*)out, n_samples, ld_in, ld_out));
    }
    void reset() { detail::check(zg_state_reset(plan_.get())); }
    void set_param(int index, const float* values, int64_t n) { detail::check(zg_param_set(plan_.get(), index, values, n)); }
    zg_plan_info info() const { zg_plan_info i; detail::check(zg_plan_get_info(plan_.get(), &i)); return i; }
};

// ---- compile() and the callable it returns -----------------------------------------------------------

namespace detail {
template <class T> constexpr int dtype_of() {
    return std::is_integral<T>::value ? ZG_I32 : std::is_same<T, float>::value ? ZG_F32 : ZG_F64;
}
// The tuple one tick of expression E returns for arguments of types Args... (flowz.hpp:1193-1201: the types fall out
// of the template-expanded evaluators there; here they are E's constexpr `types` rule, checked at every tick against
// the types of the tick program the library lowered for this signature).
template <class E, class... Args>
struct tick_result {
    static constexpr auto codes = E::types(type_list<sizeof...(Args)>{dtype_of<Args>()...});
    template <size_t... Is>
    static auto make(const double* v, const int* run_time_types, std::index_sequence<Is...>) {
        if (((run_time_types[Is] != codes[Is]) || ...))
            throw error("flowz shim: the static result type of this tick differs from the one the library computed");
        return std::tuple<typename code_type<codes[Is]>::type...>{(typename code_type<codes[Is]>::type)v[Is]...};
    }
};
}  // namespace detail

template <class E>
class stateful_lambda {
    static constexpr size_t arity = (size_t)E::in, n_out = (size_t)E::out;
    std::shared_ptr<zg_graph> graph_;
    std::unique_ptr<zg_voice, void (*)(zg_voice*)> voice_{nullptr, zg_voice_destroy};
    std::vector<const float*> refs_;

    void refresh_params() {
        for (size_t i = 0; i < refs_.size(); ++i) detail::check(zg_voice_set_param(voice_.get(), (int)i, *refs_[i]));
    }

public:
    stateful_lambda(const std::string& text, std::vector<const float*> refs) : refs_(std::move(refs)) {
        zg_graph* g = nullptr;
        detail::check(zg_graph_compile(text.c_str(), &g));
        graph_.reset(g, zg_graph_destroy);
        zg_graph_info gi;
        detail::check(zg_graph_get_info(g, &gi));
        if ((size_t)gi.n_out != n_out)      // the reference's tick would return a tuple of another size than output_arity says
            throw error("this graph returns " + std::to_string(gi.n_out) + " values per tick but its output_arity is " +
                        std::to_string(n_out) + " (sequence passes surplus inputs through, flowz.hpp:996-999)");
        zg_voice* v = nullptr;
        detail::check(zg_voice_create(g, &v));
        voice_.reset(v);
    }
    // copying the callable copies its state (flowz.hpp:1206-1207)
    stateful_lambda(const stateful_lambda& o) : graph_(o.graph_), refs_(o.refs_) {
        zg_voice* v = nullptr;
        detail::check(zg_voice_clone(o.voice_.get(), &v));
        voice_.reset(v);
    }
    stateful_lambda& operator=(const stateful_lambda& o) {
        if (this != &o) { stateful_lambda t(o); std::swap(graph_, t.graph_); std::swap(voice_, t.voice_); std::swap(refs_, t.refs_); }
        return *this;
    }
    stateful_lambda(stateful_lambda&&) = default;
    stateful_lambda& operator=(stateful_lambda&&) = default;

    // one tick; with fewer than `arity` arguments returns a closure waiting for the rest (:1203-1212)
    template <class... Args, class = std::enable_if_t<sizeof...(Args) <= arity>>
    auto operator()(const Args&... args) {
        static_assert((std::is_arithmetic<Args>::value && ...), "tick arguments must be arithmetic");
        if constexpr (sizeof...(Args) == arity) {
            double in[arity ? arity : 1] = {(double)args...};
            int dt[arity ? arity : 1] = {detail::dtype_of<Args>()...};
            double out[n_out];
            int out_types[n_out];
            refresh_params();
            detail::check(zg_voice_tick(voice_.get(), in, dt, out, out_types));
            return detail::tick_result<E, Args...>::make(out, out_types, std::make_index_sequence<n_out>{});
        } else {
            return [args..., self = *this](const auto&... rest) mutable { return self(args..., rest...); };
        }
    }

    static constexpr size_t input_count = arity;
    static constexpr size_t output_count = n_out;
    const zg_graph* graph() const { return graph_.get(); }
    std::string canonical() const { return zg_graph_canonical(graph_.get()); }
    std::string dump() const { return zg_graph_dump(graph_.get()); }
    std::vector<float> state() const {
        float* s; int n;
        detail::check(zg_voice_state(voice_.get(), &s, &n));
        return std::vector<float>(s, s + n);
    }

    // B200 block evaluator for `channels` independent voices of this graph (state starts at zero).
    block_evaluator on_device(int64_t channels, int mode = ZG_MODE_EXACT, int device = 0) const {
        zg_plan_opts o;
        zg_plan_opts_default(&o);
        o.device = device; o.channels = channels; o.mode = mode;
        return on_device(o);
    }
    block_evaluator on_device(const zg_plan_opts& o) const {
        zg_plan* p = nullptr;
        detail::check(zg_plan_create(graph_.get(), &o, &p));
        block_evaluator be(p);
        for (size_t i = 0; i < refs_.size(); ++i) be.set_param((int)i, refs_[i], 1);
        return be;
    }
};

struct compile_fn {
    template <class E, class = std::enable_if_t<detail::is_expr_v<E>>>
    auto operator()(const E& e) const {
        std::vector<const float*> refs;
        std::string text = detail::to_text(e, &refs);
        return stateful_lambda<E>(text, std::move(refs));
    }
};
inline constexpr compile_fn compile{};

}  // namespace flowz
