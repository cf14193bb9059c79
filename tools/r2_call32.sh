#!/bin/bash
# round 2, GPU call 32 (1 GPU): substitutions' block solves as products with the inverted 512 x 512 diagonal blocks
mkdir -p gpurun_out/r2c32; O=gpurun_out/r2c32
timeout 900 python -m pytest tests/test_gpu_solve_mixed.py tests/test_gpu_krylov_drivers.py tests/test_gpu_kernels.py -q -x -k "chol or solve or mixed or block_inverse or substitution or cg or refine" -p no:cacheprovider > $O/pytest_solve.log 2>&1; echo "tests rc=$?" > $O/rc.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
timeout 300 $B > $O/bench_blkinv.json 2> $O/bench_blkinv.err; echo "bench rc=$?" >> $O/rc.txt
TN_TRSV_NO_BLKINV=1 timeout 300 $B > $O/bench_noblkinv.json 2> $O/bench_noblkinv.err; echo "bench noblkinv rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
