#!/bin/bash
# round 2, GPU call 21 (1 GPU): load-only prefetch in rhs_big; the mixed solve's timing (tools/chol_one.py) with and without the L2 prefetch
mkdir -p gpurun_out/r2c21; O=gpurun_out/r2c21
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "rhs" -p no:cacheprovider > $O/pytest_rhs.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 300 python tools/rhs_probe.py > $O/rhs_probe.log 2>&1; echo "rhs rc=$?" >> $O/rc.txt
timeout 300 python tools/chol_one.py 41876 mixed 3 > $O/chol_41876_mixed.log 2>&1; echo "chol rc=$?" >> $O/rc.txt
TN_SYRK_NO_PREFETCH=1 timeout 300 python tools/chol_one.py 41876 mixed 3 > $O/chol_41876_mixed_noprefetch.log 2>&1
timeout 300 python tools/chol_one.py 41876 fp64 2 > $O/chol_41876_fp64.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
TN_SYRK_NO_PREFETCH=1 timeout 300 $B > $O/bench_noprefetch.json 2> $O/bench_noprefetch.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
