#!/bin/bash
# round 2, GPU call 30 (1 GPU): syrk_tc with the hi/lo split done by the convert kernel (four bulk copies per stage, no producer work)
mkdir -p gpurun_out/r2c30; O=gpurun_out/r2c30
TN_SYRK_PRESPLIT=1 timeout 900 python -m pytest tests/test_gpu_solve_mixed.py -q -x -p no:cacheprovider > $O/pytest_solve.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 300 python tools/chol_one.py 41876 mixed 3 > $O/chol_41876_mixed.log 2>&1
TN_SYRK_PRESPLIT=1 timeout 300 python tools/chol_one.py 41876 mixed 3 > $O/chol_41876_mixed_presplit.log 2>&1; echo "chol rc=$?" >> $O/rc.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
TN_SYRK_PRESPLIT=1 timeout 300 $B > $O/bench_presplit.json 2> $O/bench_presplit.err; echo "bench rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
