#!/bin/bash
# round 2, GPU call 36 (1 GPU): final library -- smoke, GPU suite, default bench line
mkdir -p gpurun_out/r2c36; O=gpurun_out/r2c36
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" > $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" >> $O/rc.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
