#!/bin/bash
# round 2, GPU call 43 (1 GPU): factored fp64 Gram with the tiling choice (box / consecutive, first two factors exchanged when that pads less):
# probe, GPU suite, smoke, fp64-mode bench of config 3, short default bench
mkdir -p gpurun_out/r2c43; O=gpurun_out/r2c43
timeout 150 python tools/gram_f64_probe.py quick > $O/gram_f64_probe.log 2>&1; echo "probe rc=$?" > $O/rc.txt
timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
B="--steps 2 --warmup 3 --no-peaks --no-cpu-baseline"
timeout 200 python bench.py --workload cfg3 --gram-mode fp64 $B > $O/bench_cfg3_fp64_fact.json 2> $O/bench_cfg3_fp64_fact.err; echo "cfg3 fact rc=$?" >> $O/rc.txt
timeout 200 python bench.py --rows 131072 $B > $O/bench_cfg5a_131k.json 2> $O/bench_cfg5a_131k.err; echo "default131k rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
