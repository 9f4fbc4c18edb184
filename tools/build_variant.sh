#!/bin/bash
# DEVELOPMENT TOOL: builds an A/B variant of the CUDA library -> tools/variants/<name>.so (not the product; tools/ab.py --lib loads it).
#   tools/build_variant.sh <name> [nvcc flags, e.g. -DQPIC_CELL_MINB=3] [--src <alternative qpic_ctx.cu>]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
NAME=$1; shift
SRC="$ROOT/quetzalpic_b200/csrc/qpic_ctx.cu"
mkdir -p "$HERE/variants"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -o "$HERE/variants/$NAME.so" "$SRC"
echo "built tools/variants/$NAME.so ($*)"
