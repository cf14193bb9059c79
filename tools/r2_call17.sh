#!/bin/bash
# round 2, GPU call 17 (1 GPU): run kernel with one bulk copy per stage (chunk image of Z); skeleton costs again
mkdir -p gpurun_out/r2c17; O=gpurun_out/r2c17
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -k "f16" -p no:cacheprovider > $O/pytest_run.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 600 python tools/tc16_probe.py 262144 - TN_TC16_RUN=1 TN_TC16_RUN=1,TN_TC_FLUSH_ROWS=32768 TN_TC16_RUN=1,TN_TC16_DBG=3 TN_TC16_RUN=1,TN_TC16_DBG=35 TN_TC16_RUN=1,TN_TC16_DBG=43 TN_TC16_RUN=1,TN_TC16_DBG=2 > $O/tc16_run.log 2>&1; echo "run probe rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
