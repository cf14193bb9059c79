#!/bin/bash
# round 2, GPU call 29 (1 GPU): syrk_tc with bulk-copied block images instead of the cp.async ring
mkdir -p gpurun_out/r2c29; O=gpurun_out/r2c29
timeout 900 python -m pytest tests/test_gpu_solve_mixed.py tests/test_gpu_krylov_drivers.py -q -x -p no:cacheprovider > $O/pytest_solve.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 300 python tools/chol_one.py 41876 mixed 3 > $O/chol_41876_mixed.log 2>&1; echo "chol rc=$?" >> $O/rc.txt
timeout 300 python tools/chol_one.py 16384 mixed 2 > $O/chol_16384_mixed.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
TN_FACTOR_ONE_PASS=1 timeout 300 $B > $O/bench_onepass.json 2> $O/bench_onepass.err; echo "bench onepass rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
