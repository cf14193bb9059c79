#!/bin/bash
# round 2, GPU call 20 (1 GPU): syrk_tc epilogue (L2 prefetch, 16 rows in flight): solve tests, factor timing, bench; ncu of rhs_big
mkdir -p gpurun_out/r2c20; O=gpurun_out/r2c20
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -q -x -k "chol or solve or mixed or factor" -p no:cacheprovider > $O/pytest_solve.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 300 python tools/chol_one.py 41876 mixed 2 > $O/chol_41876_mixed.log 2>&1; echo "chol rc=$?" >> $O/rc.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python tools/rhs_probe.py > $O/rhs_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_big -c 1 -o $O/ncu_rhs_big python tools/rhs_probe.py > $O/ncu_rhs_big.log 2>&1
python tools/ncu_summary.py $O/ncu_rhs_big.ncu-rep > $O/ncu_rhs_big_summary.txt 2>&1
python tools/ncu_hotspots.py $O/ncu_rhs_big.ncu-rep 30 > $O/ncu_rhs_big_hotspots.txt 2>&1
rm -f $O/ncu_rhs_big.ncu-rep
echo done >> $O/rc.txt
