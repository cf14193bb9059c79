#!/bin/bash
# round 2, GPU call 34 (1 GPU): flush window 32768 against 16384 in the sweep; other BASELINE shapes with the final kernels
mkdir -p gpurun_out/r2c34; O=gpurun_out/r2c34
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
timeout 300 $B --flush-rows 32768 > $O/bench_flush32k.json 2> $O/bench_flush32k.err; echo "bench 32k rc=$?" > $O/rc.txt
timeout 300 $B --workload cfg5b > $O/bench_cfg5b.json 2> $O/bench_cfg5b.err; echo "5b rc=$?" >> $O/rc.txt
timeout 300 $B --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "3 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peaks --workload cfg2 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "2 rc=$?" >> $O/rc.txt
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peaks --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "1 rc=$?" >> $O/rc.txt
timeout 300 $B --workload cfg4a --max-iter 50 > $O/bench_cfg4a.json 2> $O/bench_cfg4a.err; echo "4a rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
