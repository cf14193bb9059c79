#!/bin/bash
# round 2, GPU call 4: suite after the sigma fix, tf32 as the bench default, flush-window A/B, DRAM traffic of the Gram kernel
mkdir -p gpurun_out/r2c4; O=gpurun_out/r2c4
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" > $O/rc.txt
timeout 600 python tools/cfg_margins.py > $O/cfg_margins.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_1M_tf32_default.json 2> $O/bench_1M_tf32_default.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --flush-rows 32768 > $O/bench_1M_tf32_flush32k.json 2> $O/bench_1M_tf32_flush32k.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg4a --max-iter 50 > $O/bench_cfg4a.json 2> $O/bench_cfg4a.err
timeout 300 python tools/tc_one.py 131072 tf32 > $O/tc_one_plain.log 2>&1 && \
  TN_TC_FLUSH_ROWS=8192 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:gram_tc_kernel -s 1 -c 1 --csv --log-file $O/ncu_gram_tf32_traffic.csv python tools/tc_one.py 131072 tf32 > $O/ncu_gram.log 2>&1
echo done >> $O/rc.txt
