#!/bin/bash
# round 2, GPU call 23 (1 GPU): 128-byte-swizzled operand layout in the V-image kernel
mkdir -p gpurun_out/r2c23; O=gpurun_out/r2c23
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -k "run_ordered" -p no:cacheprovider > $O/pytest_run.log 2>&1; echo "tests rc=$?" > $O/rc.txt
F=TN_TC_FLUSH_ROWS=16384
timeout 600 python tools/tc16_probe.py 262144 $F $F,TN_TC16_SW128=1 $F,TN_TC16_SW128=1,TN_TC16_DBG=1 $F,TN_TC16_SW128=1,TN_TC16_DBG=2 > $O/tc16_sw128.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
TC16_SHAPE=38,6,38 timeout 600 python tools/tc16_probe.py 524288 $F $F,TN_TC16_SW128=1 > $O/tc16_sw128_5b.log 2>&1
TC16_SHAPE=24,2,24 timeout 600 python tools/tc16_probe.py 1048576 $F $F,TN_TC16_SW128=1 > $O/tc16_sw128_3.log 2>&1
echo done >> $O/rc.txt
