#!/bin/bash
# round 2, GPU call 31 (1 GPU): final state -- smoke, full GPU suite, default bench line, launch list of one sweep
mkdir -p gpurun_out/r2c31; O=gpurun_out/r2c31
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" > $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" >> $O/rc.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
timeout 200 python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-e2e --no-peaks > $O/bench_262k_plain.json 2> $O/bench_262k_plain.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 25000 --csv --log-file $O/launches_cfg5a_262k.csv python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-e2e --no-peaks > $O/bench_262k_ncu.json 2> $O/bench_262k_ncu.err; echo "launch list rc=$?" >> $O/rc.txt
python tools/launch_summary.py $O/launches_cfg5a_262k.csv 40 > $O/launches_cfg5a_262k_summary.txt 2>&1
gzip -f $O/launches_cfg5a_262k.csv
echo done >> $O/rc.txt
