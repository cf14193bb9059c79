#!/bin/bash
# round 2, GPU call 19 (1 GPU): V-image kernel as the default: f16 tests, where its time goes, bench
mkdir -p gpurun_out/r2c19; O=gpurun_out/r2c19
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -x -p no:cacheprovider > $O/pytest_tc.log 2>&1; echo "tests rc=$?" > $O/rc.txt
F=TN_TC_FLUSH_ROWS=16384
timeout 600 python tools/tc16_probe.py 262144 $F $F,TN_TC16_DBG=1 $F,TN_TC16_DBG=2 $F,TN_TC16_DBG=3 $F,TN_TC16_DBG=35 $F,TN_TC16_DBG=99 $F,TN_TC16_DBG=107 $F,TN_TC16_DBG=66 > $O/tc16_vimg_dbg.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 $B --workload cfg5b > $O/bench_cfg5b.json 2> $O/bench_cfg5b.err; echo "bench 5b rc=$?" >> $O/rc.txt
timeout 300 $B --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench 3 rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
