#!/bin/bash
# Developer tool: build flasht5_b200/libb200t5_<name>.so with extra nvcc flags on the attention kernels (A/B runs on the
# GPU box pick a library with B200T5_LIB=...).   usage: tools/build_variant.sh <name> "<extra nvcc flags>"
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/flasht5_b200/csrc; OUT=$ROOT/flasht5_b200/build/$NAME; BASE=$ROOT/flasht5_b200/build
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for f in attn_fwd attn_fwd_persist attn_bwd attn_bwd_v2; do
  nvcc $FLAGS $EXTRA -c $SRC/$f.cu -o $OUT/$f.o &
done
wait
nvcc -shared -o $ROOT/flasht5_b200/libb200t5_$NAME.so $OUT/attn_fwd.o $OUT/attn_fwd_persist.o $OUT/attn_bwd.o $OUT/attn_bwd_v2.o \
  $BASE/norm_ce.o $BASE/t5_bias.o $BASE/api.o -gencode arch=compute_100a,code=sm_100a -cudart static
ls -la $ROOT/flasht5_b200/libb200t5_$NAME.so
