#!/bin/bash
# Memory-error sweep of the SIMT kernels before they go to a GPU: the kernel sources are compiled for the CPU
# interpreter (tools/cpu_emu/cpu_emu.h) with AddressSanitizer and the emulator-backed tests run against that build,
# so an out-of-bounds index that happens to read mapped memory on the CPU (and would fault or corrupt on the GPU)
# stops the test.  TEST TOOLING ONLY.
#   tools/cpu_emu/asan.sh [pytest args]      default: the kernel-parity and widening-row test files
#   LNST_SAN=undefined tools/cpu_emu/asan.sh   the same under UndefinedBehaviorSanitizer (signed overflow, bad shifts, ...)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$(dirname "$HERE")")"
SAN=${LNST_SAN:-address}
OUT=${LNST_ASAN_DIR:-/tmp/lnst_$SAN}; mkdir -p "$OUT"
EXTRA=""; RT=libasan.so
if [ "$SAN" = undefined ]; then EXTRA="-fno-sanitize-recover=undefined"; RT=libubsan.so; fi
for s in splat field render lossnet optim gather graphnet; do
  g++ -x c++ -std=c++17 -O1 -g -fsanitize=$SAN $EXTRA -fno-omit-frame-pointer -fPIC -DLNST_CPU_EMU -Wno-unknown-pragmas \
      -I "$HERE" -c "$ROOT/neural-flow-style_b200/csrc/$s.cu" -o "$OUT/$s.o"
done
g++ -shared -fsanitize=$SAN -o "$OUT/liblnst_emu.so" "$OUT"/*.o
cd "$ROOT"
LNST_EMU_LIB="$OUT/liblnst_emu.so" LD_PRELOAD="$(gcc -print-file-name=$RT)" \
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
python -m pytest -q -p no:cacheprovider -m "not gpu" ${@:-tests/test_kernel_parity.py tests/test_widen_graphnet_kernels.py tests/test_widen_resim.py tests/test_widen_style_mask.py}
