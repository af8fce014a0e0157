#!/bin/sh
# TEST INFRASTRUCTURE: builds the CPU warp emulator of the solver kernels (tests/emu/_bin/libosb_emu.so)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=${EMU_OUT:-$HERE/_bin}      # EMU_OUT + EMU_EXTRA: a second build with other macros (tests/test_emu_kernel.py)
mkdir -p "$OUT"
g++ -std=c++17 -O2 $EMU_EXTRA -fPIC -shared -w -I"$ROOT/include" -I"$ROOT/opensot_b200/csrc" -I"$HERE" \
    "$HERE/emu_driver.cpp" "$ROOT/opensot_b200/csrc/sot_plan.cpp" "$ROOT/opensot_b200/csrc/osb_layout.c" \
    -o "$OUT/libosb_emu.so"
# the C ABI subset the C++ host layer calls, served by the emulator (tests/test_host_cpp_emu.py)
[ -n "$EMU_NO_ABI" ] && exit 0
mkdir -p "$OUT/abi"
g++ -std=c++17 -O2 $EMU_EXTRA -fPIC -shared -w -I"$ROOT/include" -I"$ROOT/opensot_b200/csrc" -I"$HERE" \
    "$HERE/emu_driver.cpp" "$HERE/emu_abi.cpp" "$ROOT/opensot_b200/csrc/sot_plan.cpp" "$ROOT/opensot_b200/csrc/osb_layout.c" \
    -Wl,-soname,libopensot_b200.so -o "$OUT/abi/libopensot_b200.so"
