mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/bwd_check.py > $O/bwd_check.log 2>&1; echo "bwd_check rc=$?"
cat $O/bwd_check.log | cut -c1-400
for m in bias nobias; do B200T5_LIB=$PWD/flasht5_b200/libb200t5_hl_bwdtiming.so timeout 100 python tools/bwd_timeline.py $m > $O/bwd3_timeline_$m.txt 2>&1; done
head -60 $O/bwd3_timeline_bias.txt
