// Minimal stand-in for <benchmark/benchmark.h> (Google Benchmark is not installed here).
// Only what /root/reference/test/benchmark.cpp needs to *compile*; nothing is ever run through it.
// Oracle-side build helper, never shipped.
#pragma once
namespace benchmark {
struct State {
    bool KeepRunning() { return false; }
};
}  // namespace benchmark
#define BENCHMARK(fn) static_assert(true, "")
#define BENCHMARK_MAIN() static_assert(true, "")
