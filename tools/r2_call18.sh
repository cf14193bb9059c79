#!/bin/bash
# round 2, GPU call 18 (1 GPU): fp16 Gram with pre-synthesised V images and pa rows (gram_tc16_vimg_kernel)
mkdir -p gpurun_out/r2c18; O=gpurun_out/r2c18
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -k "run_ordered" -p no:cacheprovider > $O/pytest_run.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 600 python tools/tc16_probe.py 262144 - TN_TC16_RUN=1 TN_TC16_VIMG=1 TN_TC16_VIMG=1,TN_TC_FLUSH_ROWS=32768 TN_TC16_VIMG=1,TN_TC_FLUSH_ROWS=16384 > $O/tc16_vimg.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
TC16_SHAPE=38,6,38 timeout 600 python tools/tc16_probe.py 524288 - TN_TC16_VIMG=1 > $O/tc16_vimg_5b.log 2>&1
TC16_SHAPE=24,2,24 timeout 600 python tools/tc16_probe.py 1048576 - TN_TC16_VIMG=1 > $O/tc16_vimg_3.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
TN_TC16_VIMG=1 timeout 300 $B > $O/bench_vimg.json 2> $O/bench_vimg.err; echo "bench vimg rc=$?" >> $O/rc.txt
TN_TC16_VIMG=1 timeout 300 $B --flush-rows 16384 > $O/bench_vimg_flush16k.json 2> $O/bench_vimg_flush16k.err; echo "bench vimg 16k rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
