#!/bin/bash
# round 2, GPU call 3: whole GPU suite with the forward-error criterion, wrappers on hardware, single-pass TF32 as preconditioner, a full default bench line
mkdir -p gpurun_out/r2c3; O=gpurun_out/r2c3
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" > $O/rc.txt
timeout 600 python tools/cfg_margins.py > $O/cfg_margins.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --gram-mode tf32 > $O/bench_1M_tf32.json 2> $O/bench_1M_tf32.err
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_1M_default.json 2> $O/bench_1M_default.err
echo done >> $O/rc.txt
