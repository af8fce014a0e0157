#!/bin/sh
# TEST INFRASTRUCTURE: builds the CPU warp emulator of the solver kernels (tests/emu/_bin/libosb_emu.so)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
mkdir -p "$HERE/_bin"
g++ -std=c++17 -O2 $EMU_EXTRA -fPIC -shared -w -I"$ROOT/include" -I"$ROOT/opensot_b200/csrc" -I"$HERE" \
    "$HERE/emu_driver.cpp" "$ROOT/opensot_b200/csrc/sot_plan.cpp" "$ROOT/opensot_b200/csrc/osb_layout.c" \
    -o "$HERE/_bin/libosb_emu.so"
# the C ABI subset the C++ host layer calls, served by the emulator (tests/test_host_cpp_emu.py)
mkdir -p "$HERE/_bin/abi"
g++ -std=c++17 -O2 $EMU_EXTRA -fPIC -shared -w -I"$ROOT/include" -I"$ROOT/opensot_b200/csrc" -I"$HERE" \
    "$HERE/emu_driver.cpp" "$HERE/emu_abi.cpp" "$ROOT/opensot_b200/csrc/sot_plan.cpp" "$ROOT/opensot_b200/csrc/osb_layout.c" \
    -Wl,-soname,libopensot_b200.so -o "$HERE/_bin/abi/libopensot_b200.so"
