#!/bin/bash
# round 2, GPU call 2: the Krylov drivers and the exact refinement on hardware
mkdir -p gpurun_out/r2c2; O=gpurun_out/r2c2
timeout 900 python -m pytest tests/test_gpu_krylov_drivers.py -q -rA -x -p no:cacheprovider > $O/pytest_drivers.log 2>&1; echo "drivers rc=$?" > $O/rc.txt
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --deselect tests/test_gpu_krylov_drivers.py > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" >> $O/rc.txt
timeout 600 python tools/cfg_margins.py > $O/cfg_margins.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --rows 131072 --no-cpu-baseline > $O/bench_131k.json 2> $O/bench_131k.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_1M.json 2> $O/bench_1M.err
TN_TC_FLUSH_ROWS=2048 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_1M_flush2048.json 2> $O/bench_1M_flush2048.err
echo done >> $O/rc.txt
