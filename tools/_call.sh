timeout 600 python tools/bwd_check.py > gpurun_out/bwd_check.log 2>&1; echo "rc=$?"; grep -E "PASS|FAIL|timing" gpurun_out/bwd_check.log | cut -c1-200
