#!/bin/bash
# Developer tool: build the variant libraries of the prepared experiments, ~10 s each.
# Each library is ~19 MB and travels with every gpurun snapshot: delete the ones a call does not need (and all of them
# before the round ends -- the driver ships the tree as it is).
set -e
cd "$(dirname "$0")/.."
python -m flasht5_b200.build > /dev/null
for k in 1 2 3 4; do tools/build_variant.sh --headline hl_poly$k "-DB200T5_EXP2_POLY=$k"; done
tools/build_variant.sh --headline hl_fhadd "-DB200T5_BIAS_FHADD=1"
tools/build_variant.sh --headline hl_fhadd_poly2 "-DB200T5_BIAS_FHADD=1 -DB200T5_EXP2_POLY=2"
tools/build_variant.sh --headline hl_timing "-DB200T5_FWD_TIMING"
tools/build_variant.sh --headline hl_timing_stagger "-DB200T5_FWD_TIMING -DB200T5_PERSIST_STAGGER_NS=700"
tools/build_variant.sh --headline hl_stagger "-DB200T5_PERSIST_STAGGER_NS=700"
tools/build_variant.sh --headline hl_bwdpp "-DB200T5_BWD_PINGPONG=1"
tools/build_variant.sh --headline hl_bwdpp_poly2 "-DB200T5_BWD_PINGPONG=1 -DB200T5_EXP2_POLY=2"
ls -la flasht5_b200/libb200t5_*.so
