// A flowz user's program, unchanged in spelling, built against the B200-native library instead of the
// reference header (andre-bergner/zignal flowz/flowz.hpp).  Used by tests/test_cpp_dropin.py:
//   ./dropin          host only: compile() + operator() ticks (BASELINE configs[0], runs without a GPU)
//   ./dropin gpu      additionally: the same graphs as blocks on the B200 through on_device(), which must
//                     reproduce the per-sample ticks bit for bit (EXACT mode)
#include <flowz/flowz.hpp>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>

static int failures = 0;
#define CHECK(...)                                                           \
    do {                                                                     \
        if (!(__VA_ARGS__)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #__VA_ARGS__); ++failures; } \
    } while (0)

static float noise(unsigned& s) {                       // deterministic input in [-1, 1)
    s = s * 1664525u + 1013904223u;
    return (float)((s >> 8) & 0xffffff) / 8388608.0f - 1.0f;
}

int main(int argc, char** argv) {
    using namespace flowz;
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;

    // ---- BASELINE configs[0]: mono one-pole low-pass y = x + a*y1 over 1024 samples (flowz/README.md) ----
    auto one_pole = compile(~(_2 + 0.9f * _1[_1]));
    std::vector<float> x(1024), y_host(1024);
    unsigned seed = 1;
    for (auto& v : x) v = noise(seed);
    float y1 = 0.f;
    for (int t = 0; t < 1024; ++t) {
        y_host[t] = std::get<0>(one_pole(x[t]));
        const float want = x[t] + 0.9f * y1;             // the same arithmetic by hand
        CHECK(y_host[t] == want);
        y1 = want;
    }

    // ---- the reference's benchmark graph: direct-form-1 biquad, fwd |= bwd (test/benchmark.cpp:18-33) ----
    const float b0 = 0.2f, b1 = 0.4f, b2 = 0.2f, a1 = 0.3f, a2 = -0.1f;
    auto fwd = b0 * _1 + b1 * _1[_1] + b2 * _1[_2];
    auto bwd = ~(_2 + a1 * _1[_1] + a2 * _1[_2]);
    auto biquad = compile(fwd |= bwd);
    std::vector<float> yb_host(1024);
    {
        float x1 = 0, x2 = 0, v1 = 0, v2 = 0;
        for (int t = 0; t < 1024; ++t) {
            yb_host[t] = std::get<0>(biquad(x[t]));
            const float v = (b0 * x[t] + b1 * x1) + b2 * x2;
            const float w = (v + a1 * v1) + a2 * v2;
            CHECK(yb_host[t] == w);
            x2 = x1; x1 = x[t]; v2 = v1; v1 = w;
        }
    }

    // ---- std::ref parameters (flowz/README.md:42-63) and state copy (flowz.hpp:1206-1207) ----
    float gain = 0.5f;
    auto amp = compile(std::ref(gain) * _1);
    CHECK(std::get<0>(amp(2.0f)) == 1.0f);
    gain = 3.0f;
    CHECK(std::get<0>(amp(2.0f)) == 6.0f);
    auto fork = one_pole;                                 // continues from the same state, independently
    CHECK(std::get<0>(fork(0.25f)) == std::get<0>(one_pole(0.25f)));

    // ---- currying (flowz.hpp:1203-1212): with fewer arguments than inputs, operator() returns a closure that holds
    //      the given arguments and a COPY of the callable (`expr = *this`, state included); the closure is mutable, so
    //      its copy advances from call to call while the original never moves ----
    {
        auto two = compile(~(_2 + _3 + 0.5f * _1[_1]));          // two inputs: y = a + b + 0.5 * y[-1]
        auto direct = two;                                       // an independent voice, same (zero) state
        const std::vector<float> zero_state = two.state();
        auto half = two(1.0f);                                   // waits for the second argument
        const float d1 = std::get<0>(direct(1.0f, 2.0f)), d2 = std::get<0>(direct(1.0f, 2.0f));
        CHECK(d1 == 3.0f && d2 == 4.5f);
        CHECK(std::get<0>(half(2.0f)) == d1);                    // f(x1)(x2) == f(x1, x2)
        CHECK(std::get<0>(half(2.0f)) == d2);                    // the closure's own copy carries its state on
        CHECK(two.state() == zero_state);                        // ... and the original has not ticked at all
        CHECK(std::get<0>(two(1.0f, 2.0f)) == d1);               // so it still starts from zero
        auto later = two(1.0f);                                  // a closure made now copies the advanced state
        CHECK(std::get<0>(later(2.0f)) == d2);
        auto none = two();                                       // no arguments at all: waits for both
        CHECK(std::get<0>(none(1.0f, 2.0f)) == d2);
        auto three = compile(_1 + 2.0f * _2 + 4.0f * _3);        // closures curry again: f(a)(b)(c)
        CHECK(std::get<0>(three(1.0f)(1.0f)(1.0f)) == 7.0f && std::get<0>(three(1.0f, 1.0f)(1.0f)) == 7.0f);
    }

    // ---- the tuple a tick returns has the reference's per-wire C++ types (flowz.hpp:1193-1201): the reference's own
    //      integer ticks (test/tests.cpp:88-134) in its spelling, plus the types they imply ----
    {
        using std::tuple;
        using std::is_same;
        auto wp = compile(_1 |= _2);                             // wire-around (test/tests.cpp:88-95)
        auto wpr = wp(2, 1337);
        CHECK(std::make_tuple(1337) == wpr);
        static_assert(is_same<decltype(wpr), tuple<int>>::value, "int in, int out");
        auto ws = compile((_1, _1) |= _1);                       // :97-103
        auto wsr = ws(1337);
        CHECK(std::make_tuple(1337, 1337) == wsr);
        static_assert(is_same<decltype(wsr), tuple<int, int>>::value, "");
        auto identity = compile(_1);                             // :111-114
        CHECK(std::make_tuple(1337) == identity(1337) && std::make_tuple(42) == identity(42));
        static_assert(is_same<decltype(identity(42)), tuple<int>>::value && is_same<decltype(identity(4.2)), tuple<double>>::value, "");
        auto unit_delay = compile(_1[_1]);                       // :117-121: the delayed value comes out of the float state
        CHECK(std::make_tuple(0) == unit_delay(1337) && std::make_tuple(1337) == unit_delay(42) && std::make_tuple(42) == unit_delay(17));
        static_assert(is_same<decltype(unit_delay(17)), tuple<float>>::value, "");
        auto differentiator = compile(_1 - _1[_1]);              // :124-128
        CHECK(std::make_tuple(1337) == differentiator(1337) && std::make_tuple(42 - 1337) == differentiator(42) &&
              std::make_tuple(17 - 42) == differentiator(17));
        static_assert(is_same<decltype(differentiator(17)), tuple<float>>::value, "int - float");
        auto integrator = compile(~(_1[_1] + _2));               // :131-135
        CHECK(std::make_tuple(1337) == integrator(1337) && std::make_tuple(1337 + 42) == integrator(42) &&
              std::make_tuple(1337 + 42 + 17) == integrator(17));
        static_assert(is_same<decltype(integrator(17)), tuple<float>>::value && is_same<decltype(integrator(1.5)), tuple<double>>::value, "");
        // mixed wires: every output has its own type (test/tests.cpp:206-211 for the same expressions as ResultType)
        auto mixed = compile((_1, _1 * 1.0));
        static_assert(is_same<decltype(mixed(1.0f)), tuple<float, double>>::value, "");
        static_assert(is_same<decltype(mixed(1)), tuple<int, double>>::value, "");
        CHECK(mixed(3) == std::make_tuple(3, 3.0));
        auto swapped = compile((_1 * 1.0, _1) |= (_2, _1));
        static_assert(is_same<decltype(swapped(1.0f)), tuple<float, double>>::value, "");
        auto sums = compile(_1 + _2);
        static_assert(is_same<decltype(sums(1, 2)), tuple<int>>::value && is_same<decltype(sums(1, 2.f)), tuple<float>>::value &&
                      is_same<decltype(sums(1., 2.f)), tuple<double>>::value, "usual arithmetic conversions");
        CHECK(sums(7, 2) == std::make_tuple(9) && std::get<0>(compile(_1 / _2)(7, 2)) == 3);    // int / int stays an int division
        static_assert(is_same<decltype(one_pole(1.0f)), tuple<float>>::value && is_same<decltype(one_pole(1.0)), tuple<double>>::value &&
                      is_same<decltype(one_pole(1)), tuple<float>>::value, "feedback: the type of what is fed back");
        static_assert(is_same<decltype(biquad(1.0f)), tuple<float>>::value, "");
    }

    // ---- expr[_n], a spelling the reference plans (TODO.md:51-52): (_1+_2)[_1] == _1+_2 |= _1[_1] ----
    {
        auto a = compile((_1 + _2)[_1]);
        auto b = compile(_1 + _2 |= _1[_1]);
        for (int t = 0; t < 8; ++t) CHECK(a(float(t), 2.f) == b(float(t), 2.f));
        CHECK(same_expr((_1 + _2)[_1], _1 + _2 |= _1[_1]));
        CHECK(same_expr((0.5f * _1)[-3], 0.5f * _1 |= _1[_3]));
    }

    // ---- nested feedback, which the reference cannot split (TODO.md:11-27, disabled test/tests.cpp:59):
    //      s[t] = y + y + 1 with y = s[t-1]  ->  y = 0, 1, 3, 7, 15 ----
    {
        auto nested = compile(~~(_1 + _2 + 1 |= _1[_1]));
        const float want[5] = {0.f, 1.f, 3.f, 7.f, 15.f};
        for (int t = 0; t < 5; ++t) CHECK(std::get<0>(nested()) == want[t]);
        bool threw = false;
        try { (void)compile(~(_1 + _2)); } catch (const std::exception& e) { threw = std::strstr(e.what(), "without a delay") != nullptr; }
        CHECK(threw);
    }

    // ---- ResultType (flowz.hpp:515-644), the reference's test_result_type_transform (test/tests.cpp:182-232) in the
    //      reference's spelling; `expect_type(T{}, expr)` becomes r(expr, x).is<T>() because types are run-time data ----
    {
        using namespace flowz::transforms;
        using std::tuple;
        using cplx = std::complex<float>;
        ResultType r;
        tuple<float> xin;
        CHECK(r(_1, xin).is<float>());
        CHECK(r(_1 * 1.0, xin).is<double>());
        CHECK(r(_1 |= _1, xin).is<tuple<float>>());
        CHECK(r(_1 * 1.0 |= _1, xin).is<tuple<double>>());
        CHECK(r(_1 |= 1.0 * _1, xin).is<tuple<double>>());
        CHECK(r(_1 |= cplx{1, 0} * _1, xin).is<tuple<cplx>>());
        CHECK(r(2 * _1 |= _1 * cplx{1, 0} |= _1, xin).is<tuple<cplx>>());
        CHECK(r((_1, _1), xin).is<tuple<float, float>>());
        CHECK(r((_1, _1 * 1.0), xin).is<tuple<float, double>>());
        CHECK(r((_1 * 1.0, _1), xin).is<tuple<double, float>>());
        CHECK(r((_1 * 1.0, _1) |= (_2, _1), xin).is<tuple<float, double>>());
        CHECK(r((_1, 1.0 * _1) |= (_1 | _1), xin).is<tuple<float, double>>());
        CHECK(r(_1[_1], xin).is<float>());
        CHECK(r((_1[_1], 1.0 * _1) |= _2[_1], xin).is<tuple<double>>());
        CHECK(r(~(_1[_1] + _2), xin).is<tuple<float>>());
        CHECK(r(~(1.0 * _1[_1] + _2), xin).is<tuple<double>>());
        CHECK(r(~(_1[_1] + 1.0), xin).is<tuple<double>>());
        make_canonical c;
        CHECK(r(c(~(_1[_1] + 1.0 * _2)), xin).is<tuple<double>>());
        CHECK(!r(~(_1[_1] + _2), xin).is<tuple<double>>());
    }

    if (gpu) {
        // the same graphs for 96 voices at once; every voice gets the same input, so every row must equal
        // the host ticks above bit for bit
        const int64_t C = 96, T = 1024;
        std::vector<float> in((size_t)C * T), out((size_t)C * T);
        for (int64_t c = 0; c < C; ++c) std::memcpy(&in[c * T], x.data(), T * sizeof(float));
        const float* ip[1] = {in.data()};
        float* op[1] = {out.data()};
        {
            auto dev = compile(~(_2 + 0.9f * _1[_1])).on_device(C);
            dev.process_host(ip, op, 512, T, T);                       // two blocks: state carries over
            const float* ip2[1] = {in.data() + 512};
            float* op2[1] = {out.data() + 512};
            dev.process_host(ip2, op2, 512, T, T);
            for (int64_t c = 0; c < C; ++c) CHECK(std::memcmp(&out[c * T], y_host.data(), T * sizeof(float)) == 0);
        }
        {
            auto dev = compile(fwd |= bwd).on_device(C);
            dev.process_host(ip, op, T, T, T);
            for (int64_t c = 0; c < C; ++c) CHECK(std::memcmp(&out[c * T], yb_host.data(), T * sizeof(float)) == 0);
            CHECK(std::strstr(dev.info().kernel, "zg_biquad_df1") != nullptr);   // recognised: prebuilt kernel
        }
    } else {
        // without a device the block evaluator must say so, not fall back to a CPU loop
        bool threw = false;
        try { (void)one_pole.on_device(8); } catch (const std::exception& e) { threw = std::strstr(e.what(), "no CPU fallback") != nullptr; }
        if (!threw) std::printf("note: on_device() did not fail -- is a GPU visible?\n");
    }

    std::printf(failures ? "dropin: %d FAILURES\n" : "dropin ok (%d failures)\n", failures);
    return failures ? 1 : 0;
}
