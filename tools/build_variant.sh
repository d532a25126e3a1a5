#!/bin/bash
# Developer tool: build flasht5_b200/libb200t5_<name>.so with extra nvcc flags on the attention kernels; A/B runs on the
# GPU box pick a library with B200T5_LIB=... (tools/lib_ab_check.py, tools/gpu_perf.py).
#   usage: tools/build_variant.sh [--headline] <name> "<extra nvcc flags>"
# --headline: only the D = 64 / bf16 instantiations (-DB200T5_HEADLINE_ONLY) and the stock D = 128 backward object:
#             ~20 s to build and a ~5 MB library instead of ~75 s / 25 MB.
# Known switches: -DB200T5_EXP2_POLY=K  -DB200T5_BIAS_FHADD=1  -DB200T5_BWD_PINGPONG=1  -DB200T5_PERSIST_STAGGER_NS=N
#                 -DB200T5_PRODUCER_SLEEP_NS=N  -DB200T5_FWD_TIMING  -DB200T5_BWD_TIMING
set -e
HEADLINE=0
if [ "$1" == "--headline" ]; then HEADLINE=1; shift; fi
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/flasht5_b200/csrc; OUT=$ROOT/flasht5_b200/build/$NAME; BASE=$ROOT/flasht5_b200/build
[ -f $BASE/api.o ] || python -m flasht5_b200.build > /dev/null
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
FILES="attn_fwd attn_fwd_persist attn_fwd_pingpong attn_bwd_v2"
if [ $HEADLINE == 1 ]; then EXTRA="$EXTRA -DB200T5_HEADLINE_ONLY"; cp $BASE/attn_bwd.o $OUT/attn_bwd.o; else FILES="$FILES attn_bwd"; fi
for f in $FILES; do
  nvcc $FLAGS $EXTRA -c $SRC/$f.cu -o $OUT/$f.o &
done
wait
nvcc -shared -o $ROOT/flasht5_b200/libb200t5_$NAME.so $OUT/attn_fwd.o $OUT/attn_fwd_persist.o $OUT/attn_fwd_pingpong.o $OUT/attn_bwd.o $OUT/attn_bwd_v2.o \
  $BASE/norm_ce.o $BASE/t5_bias.o $BASE/adamw.o $BASE/api.o -gencode arch=compute_100a,code=sm_100a -cudart static
ls -la $ROOT/flasht5_b200/libb200t5_$NAME.so
