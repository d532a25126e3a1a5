#!/bin/bash
# Developer tool: build flasht5_b200/libb200t5_<name>.so with extra nvcc flags on the attention kernels; A/B runs on the
# GPU box pick a library with B200T5_LIB=... (tools/bwd_check.py, tools/gpu_perf.py).
#   usage: tools/build_variant.sh [--headline] <name> "<extra nvcc flags>"
# --headline: only the D = 64 / bf16 instantiations (-DB200T5_HEADLINE_ONLY): ~20 s to build and a small library.
# Known switches: -DB200T5_FWD_TIMING  -DB200T5_BWD_TIMING  -DB200T5_DEBUG_DEADLOCK  -DB200T5_RPE_SKIP_LEVEL=0|1|2
#   backward ablations (wrong results by construction): -DB200T5_DBG_SKIP_EXP  _TRUNC_PACK  _SKIP_MATH  _SKIP_DS_TMA  _SKIP_DQ_TMA
set -e
HEADLINE=0
if [ "$1" == "--headline" ]; then HEADLINE=1; shift; fi
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/flasht5_b200/csrc; OUT=$ROOT/flasht5_b200/build/$NAME; BASE=$ROOT/flasht5_b200/build
[ -f $BASE/api.o ] || python -m flasht5_b200.build > /dev/null
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
FILES="attn_fwd attn_bwd_v3 api"
if [ $HEADLINE == 1 ]; then EXTRA="$EXTRA -DB200T5_HEADLINE_ONLY"; fi
for f in $FILES; do
  nvcc $FLAGS $EXTRA -c $SRC/$f.cu -o $OUT/$f.o &
done
wait
nvcc -shared -o $ROOT/flasht5_b200/libb200t5_$NAME.so $OUT/attn_fwd.o $BASE/attn_bwd.o $OUT/attn_bwd_v3.o \
  $BASE/norm_ce.o $BASE/t5_bias.o $BASE/adamw.o $OUT/api.o -gencode arch=compute_100a,code=sm_100a -cudart static
ls -la $ROOT/flasht5_b200/libb200t5_$NAME.so
