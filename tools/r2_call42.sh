#!/bin/bash
# round 2, GPU call 42 (1 GPU): factored fp64 Gram with the round-aware row split and the two-warp feature expansion: probe, ncu captures,
# fp64-mode bench lines (config 3 against the unfactored kernel, config 5b)
mkdir -p gpurun_out/r2c42; O=gpurun_out/r2c42
timeout 200 python tools/gram_f64_probe.py > $O/gram_f64_probe.log 2>&1; echo "probe rc=$?" > $O/rc.txt
B="--steps 2 --warmup 3 --no-peaks --no-cpu-baseline --gram-mode fp64"
timeout 200 python bench.py --workload cfg3 $B > $O/bench_cfg3_fp64_fact.json 2> $O/bench_cfg3_fp64_fact.err; echo "cfg3 fact rc=$?" >> $O/rc.txt
TN_GRAM_F64_UNFACTORED=1 timeout 200 python bench.py --workload cfg3 $B > $O/bench_cfg3_fp64_unfactored.json 2> $O/bench_cfg3_fp64_unfactored.err; echo "cfg3 old rc=$?" >> $O/rc.txt
timeout 200 python bench.py --workload cfg5b --rows 131072 $B > $O/bench_cfg5b_fp64_fact.json 2> $O/bench_cfg5b_fp64_fact.err; echo "cfg5b fact rc=$?" >> $O/rc.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gram_f64_fact -s 1 -c 1 -o $O/ncu_f64fact_cfg5a python tools/tc_one.py 8192 fp64 38,29,38 > $O/ncu_cfg5a.log 2>&1; echo "ncu5a rc=$?" >> $O/rc.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gram_f64_fact -s 1 -c 1 -o $O/ncu_f64fact_cfg3 python tools/tc_one.py 65536 fp64 24,2,24 > $O/ncu_cfg3.log 2>&1; echo "ncu3 rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
