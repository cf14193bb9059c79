#!/bin/bash
# round 2, GPU call 22 (1 GPU): CTA-pair variant of the V-image fp16 Gram kernel
mkdir -p gpurun_out/r2c22; O=gpurun_out/r2c22
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -k "run_ordered" -p no:cacheprovider > $O/pytest_run.log 2>&1; echo "tests rc=$?" > $O/rc.txt
F=TN_TC_FLUSH_ROWS=16384
timeout 600 python tools/tc16_probe.py 262144 $F $F,TN_TC16_VPAIR=1 > $O/tc16_vpair.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
TC16_SHAPE=38,6,38 timeout 600 python tools/tc16_probe.py 524288 $F $F,TN_TC16_VPAIR=1 > $O/tc16_vpair_5b.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
TN_TC16_VPAIR=1 timeout 300 $B > $O/bench_vpair.json 2> $O/bench_vpair.err; echo "bench rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
