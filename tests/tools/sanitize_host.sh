#!/bin/bash
# Host code under AddressSanitizer + UndefinedBehaviorSanitizer (no GPU): analysis with the complex element size, the real
# embedding, the plan builder, the CPU interpreter of the schedule (blocked path: nb 8 / outer 16, no small fronts), complex LU
# with row phases, and the forked parallel ordering.  Expected output: one "rc 0" line per case and no sanitizer report.
#   bash tests/tools/sanitize_host.sh
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
CSRC="$ROOT/sparse-matrix-factorization-library_b200/csrc"
OUT="${TMPDIR:-/tmp}/sfb200_sanitize"
mkdir -p "$OUT"
make -C "$ROOT/oracle" _ref/xstat.o > /dev/null
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -w -I"$CSRC" "$ROOT/tests/tools/sanitize_host_main.cpp" \
    "$ROOT/oracle/plan_host_exec.cpp" "$ROOT/oracle/plan_race_check.cpp" "$CSRC/plan.cpp" "$CSRC/symbolic.cpp" \
    "$CSRC/parallel_order.cpp" "$CSRC/complex_embed.cpp" "$ROOT/oracle/_ref/xstat.o" \
    /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -pthread -o "$OUT/run"
ASAN_OPTIONS=detect_leaks=0 "$OUT/run"
# the MatrixMarket / binary reader on 300 000 mutated inputs (header-only, no GPU library needed)
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -w -I"$ROOT/sparse-matrix-factorization-library_b200/dropin" \
    "$ROOT/tests/tools/fuzz_reader.cpp" -pthread -o "$OUT/fuzz_reader"
ASAN_OPTIONS=detect_leaks=0 "$OUT/fuzz_reader" 300000
# a foreign supernodal structure damaged in a few places must be refused or digested, never indexed out of bounds
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -w -I"$CSRC" "$ROOT/tests/tools/fuzz_wrap.cpp" \
    "$CSRC/capi_symbolic.cpp" "$CSRC/plan.cpp" "$CSRC/symbolic.cpp" "$CSRC/parallel_order.cpp" "$CSRC/complex_embed.cpp" \
    "$ROOT/oracle/_ref/xstat.o" /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -pthread -o "$OUT/fuzz_wrap"
ASAN_OPTIONS=detect_leaks=0 "$OUT/fuzz_wrap" 5000
