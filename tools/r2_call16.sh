#!/bin/bash
# round 2, GPU call 16 (1 GPU): which part of the fp16 Gram kernel's per-stage skeleton costs the time
mkdir -p gpurun_out/r2c16; O=gpurun_out/r2c16
timeout 600 python tools/tc16_probe.py 262144 TN_TC16_RUN=1 TN_TC16_RUN=1,TN_TC16_DBG=3 TN_TC16_RUN=1,TN_TC16_DBG=35 TN_TC16_RUN=1,TN_TC16_DBG=39 TN_TC16_RUN=1,TN_TC16_DBG=43 TN_TC16_RUN=1,TN_TC16_DBG=51 TN_TC16_RUN=1,TN_TC16_DBG=63 > $O/tc16_dbg.log 2>&1; echo "rc=$?" > $O/rc.txt
echo done >> $O/rc.txt
