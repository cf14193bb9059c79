#!/bin/bash
# round 2, GPU call 13 (1 GPU): register-blocked rhs kernel, 16-warp run-ordered Gram producers, one-pass TF32 factor; short bench A/B runs
mkdir -p gpurun_out/r2c13; O=gpurun_out/r2c13
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gram_tc.py tests/test_gpu_krylov_drivers.py -q -rA -x -k "rhs or run_ordered or krylov or cg or refine" -p no:cacheprovider > $O/pytest_new.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 300 python tools/rhs_probe.py > $O/rhs_probe.log 2>&1; echo "rhs probe rc=$?" >> $O/rc.txt
timeout 600 python tools/tc16_probe.py 262144 - TN_TC16_RUN=1 TN_TC16_RUN=2 TN_TC16_RUN=2,TN_TC_FLUSH_ROWS=32768 > $O/tc16_run.log 2>&1; echo "run probe rc=$?" >> $O/rc.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
timeout 300 $B > $O/bench_a_default.json 2> $O/bench_a.err; echo "bench a rc=$?" >> $O/rc.txt
TN_FACTOR_ONE_PASS=1 timeout 300 $B > $O/bench_b_onepass.json 2> $O/bench_b.err; echo "bench b rc=$?" >> $O/rc.txt
TN_FACTOR_ONE_PASS=1 TN_TC16_RUN=1 timeout 300 $B --flush-rows 32768 > $O/bench_c_run1_flush32k_onepass.json 2> $O/bench_c.err; echo "bench c rc=$?" >> $O/rc.txt
TN_FACTOR_ONE_PASS=1 TN_TC16_RUN=2 timeout 300 $B --flush-rows 32768 > $O/bench_d_run2_flush32k_onepass.json 2> $O/bench_d.err; echo "bench d rc=$?" >> $O/rc.txt
TN_TC16_RUN=1 timeout 300 $B > $O/bench_e_run1.json 2> $O/bench_e.err; echo "bench e rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
