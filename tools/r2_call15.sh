#!/bin/bash
# round 2, GPU call 15 (1 GPU): which side bounds the fp16 Gram kernel -- producers without MMAs, MMAs without producers
mkdir -p gpurun_out/r2c15; O=gpurun_out/r2c15
timeout 600 python tools/tc16_probe.py 262144 TN_TC16_RUN=1 TN_TC16_RUN=1,TN_TC16_DBG=1 TN_TC16_RUN=1,TN_TC16_DBG=2 TN_TC16_RUN=1,TN_TC16_DBG=3 > $O/tc16_dbg.log 2>&1; echo "rc=$?" > $O/rc.txt
echo done >> $O/rc.txt
