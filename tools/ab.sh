#!/bin/bash
# Developer tool (GPU box): A/B several builds of the library in ONE call (boxes differ by ~10%).  usage: tools/ab.sh shape lib1 lib2 ...
SHAPE=$1; shift
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
for rep in 1 2; do
for LIB in "$@"; do
  echo "== $LIB (rep $rep)"
  B200T5_LIB=$PWD/flasht5_b200/$LIB tools/launch_times.sh $SHAPE ab_$(basename $LIB .so) | grep -E "attn_fwd_kernel|attn_bwd_kernel"
done
done
