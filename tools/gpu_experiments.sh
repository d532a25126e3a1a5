#!/bin/bash
# Developer tool (GPU box): one pass over the prepared experiments of DESIGN.md section 9.  Build the variant libraries
# first (tools/build_all_variants.sh, in the build container); everything is written under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_experiments.sh'
mkdir -p gpurun_out
O=gpurun_out
echo "== gated GPU tests (AdamWScale, relative-position backward level 2)" | tee $O/exp_summary.txt
B200T5_ADAMW_GPU=1 B200T5_RPE_SKIP2_GPU=1 B200T5_RMSNORM_PREFETCH_GPU=1 B200T5_T5BIAS_TILES_GPU=1 timeout 200 python -m pytest tests/test_adamw_scaled.py tests/test_attention_rpe.py tests/test_norm_ce_gpu.py tests/test_positional_encoding.py -m gpu -q > $O/exp_gated_tests.log 2>&1
tail -3 $O/exp_gated_tests.log | tee -a $O/exp_summary.txt
echo "== relative-position backward: skip levels" | tee -a $O/exp_summary.txt
timeout 120 python tools/rpe_skip_check.py > $O/exp_rpe_skip.log 2>&1
grep -E "summary|timing" $O/exp_rpe_skip.log | tee -a $O/exp_summary.txt
echo "== forward kernels: base / persistent / ping-pong (bit-identity, then times)" | tee -a $O/exp_summary.txt
timeout 200 python tools/fwd_persist_check.py > $O/exp_fwd_kernels.log 2>&1
grep -E "equal_summary|timing|error" $O/exp_fwd_kernels.log | cut -c1-400 | tee -a $O/exp_summary.txt
echo "== variant libraries against the stock one" | tee -a $O/exp_summary.txt
timeout 100 python tools/lib_ab_check.py --save /tmp/ab_base.pt > $O/exp_ab_base.log 2>&1
grep '"lib"' $O/exp_ab_base.log | tee -a $O/exp_summary.txt
for lib in flasht5_b200/libb200t5_hl_*.so; do
  [ -f "$lib" ] || continue
  case "$lib" in *timing*) continue;; esac
  n=$(basename $lib .so)
  B200T5_LIB=$PWD/$lib timeout 100 python tools/lib_ab_check.py --compare /tmp/ab_base.pt --tol > $O/exp_ab_$n.log 2>&1
  grep -E 'AB_CHECK|"lib"' $O/exp_ab_$n.log | cut -c1-300 | tee -a $O/exp_summary.txt
  B200T5_FWD_PERSIST=1 B200T5_LIB=$PWD/$lib timeout 100 python tools/lib_ab_check.py --compare /tmp/ab_base.pt --tol > $O/exp_ab_${n}_persist.log 2>&1
  grep -E '"lib"' $O/exp_ab_${n}_persist.log | sed 's/^/persist: /' | cut -c1-300 | tee -a $O/exp_summary.txt
done
echo "== ping-pong forward timeline" | tee -a $O/exp_summary.txt
[ -f flasht5_b200/libb200t5_hl_timing.so ] && B200T5_FWD_PINGPONG=1 B200T5_LIB=$PWD/flasht5_b200/libb200t5_hl_timing.so timeout 60 python tools/fwd_timeline.py bias > $O/exp_timeline_pingpong.txt 2>&1
echo "== persistent forward timeline, with and without the stagger" | tee -a $O/exp_summary.txt
for n in hl_timing hl_timing_stagger; do
  [ -f flasht5_b200/libb200t5_$n.so ] || continue
  B200T5_FWD_PERSIST=1 B200T5_LIB=$PWD/flasht5_b200/libb200t5_$n.so timeout 60 python tools/fwd_timeline.py bias > $O/exp_timeline_persist_$n.txt 2>&1
done
echo "== AdamWScale against the HBM roof" | tee -a $O/exp_summary.txt
timeout 120 python tools/gpu_perf_adamw.py --out $O/exp_adamw.json 2>&1 | tail -1 | cut -c1-400 | tee -a $O/exp_summary.txt
timeout 120 python tools/gpu_perf_adamw.py --dtype bf16 --kahan --out $O/exp_adamw.json 2>&1 | tail -1 | cut -c1-400 | tee -a $O/exp_summary.txt
