#!/bin/bash
# The host half of the C ABI (parser, analyses, canonical split, lowering, host voice) under AddressSanitizer,
# UndefinedBehaviorSanitizer and LeakSanitizer over a few thousand random, golden and malformed expressions.
# Usage: bash tools/sanitize/run.sh        (CPU only; prints "N expressions, M valid graphs" and no sanitizer report)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${TMPDIR:-/tmp}/zg_sanitize
mkdir -p "$OUT"
R=$ROOT/zignal_b200/csrc
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -I"$ROOT/include" -I"$R" \
    "$ROOT/tools/sanitize/host_abi_driver.cpp" "$R/zg_expr.cpp" "$R/zg_ir.cpp" "$R/zg_capi.cpp" "$R/zg_codegen.cpp" "$R/zg_match.cpp" "$R/zg_scan.cpp" \
    -o "$OUT/harness"
PYTHONPATH="$ROOT/tests:$ROOT/oracle:$ROOT/tests/golden:$ROOT" python - > "$OUT/exprs.txt" <<'PY'
import random
import test_fuzz_frontend as tf
import reference_vectors as rv
rng = random.Random(5)
for _ in range(6000):
    print(tf._gen(rng, rng.randint(1, 6), rng.randint(1, 5)))
for c in rv.CANONICAL_SPLITS: print(c[0])
for c in rv.RESULT_TYPES: print(c[0])
for s in ["", "(", "_", "_0", "_1[", "_1[_0]", "~", "_1 |=", "1e999", "0x", "cplx{", "cplx{1,", "_1<-", "$", "$-1", "bfb(_1", "front(0)",
          "front(99999)", "_99999999999", "_1[_99999999999]", ")" * 10, "(" * 2000, "-" * 3000 + "_1", "~" * 3000 + "_1",
          "_1" + "[_1]" * 3000, "_1[_2000000000]", "_1 * $2000000000", "_1[_4294967297]", "$4294967296", "3000000000*_1", "_1[_1048576]", "_1 / _2", "_1[_1000000] + _1[_999999]", "_1 " + "+ _1 " * 5000, "(" * 1500 + "_1" + ")" * 1500, "_1" + " |= _1" * 1500, "_1" + " + _1[_2]" * 1500]:
    print(s)
PY
# sanitizer frames are several times larger than the product's: give the deepest accepted expressions room
ulimit -s 1000000 2>/dev/null || ulimit -s unlimited 2>/dev/null || true
ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 "$OUT/harness" < "$OUT/exprs.txt"

# the same sources under ThreadSanitizer: one graph shared by 8 threads (host_abi_threads.cpp)
g++ -std=c++17 -O1 -g -fsanitize=thread -ffp-contract=off -I"$ROOT/include" -I"$R" \
    "$ROOT/tools/sanitize/host_abi_threads.cpp" "$R/zg_expr.cpp" "$R/zg_ir.cpp" "$R/zg_capi.cpp" "$R/zg_codegen.cpp" "$R/zg_match.cpp" "$R/zg_scan.cpp" \
    -o "$OUT/tsan_harness" -lpthread
"$OUT/tsan_harness"
