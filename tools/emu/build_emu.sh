#!/bin/bash
# DEVELOPMENT TOOL: compiles the CUDA sources for the CPU fiber emulator (kernel-logic debugging without a GPU).
# Output: tools/emu/_build/libqpic_emu_debug.so -- never loaded by the quetzalpic_b200 package.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
mkdir -p "$HERE/_build"
/usr/bin/g++ -std=c++17 -O1 -g -fPIC -shared -DQPIC_EMU -I"$HERE" -I"$ROOT/quetzalpic_b200/csrc" \
    $QPIC_EMU_FLAGS -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable \
    -x c++ "$ROOT/quetzalpic_b200/csrc/qpic_ctx.cu" "$HERE/qpic_emu.cpp" \
    -o "$HERE/_build/libqpic_emu_debug.so"
echo "built $HERE/_build/libqpic_emu_debug.so"
