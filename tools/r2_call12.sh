#!/bin/bash
# round 2, GPU call 12 (1 GPU): run-ordered producers v2 (piece-per-lane, padded strides): test, A/B timing, raw ncu metrics; launch list of one bench sweep
mkdir -p gpurun_out/r2c12; O=gpurun_out/r2c12
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -k "run_ordered" -p no:cacheprovider > $O/pytest_f16.log 2>&1; echo "f16 tests rc=$?" > $O/rc.txt
timeout 600 python tools/tc16_probe.py 262144 - TN_TC16_RUN=1 TN_TC16_RUN=1,TN_TC_FLUSH_ROWS=32768 > $O/tc16_run.log 2>&1; echo "run probe rc=$?" >> $O/rc.txt
TC16_SHAPE=38,6,38 timeout 600 python tools/tc16_probe.py 524288 - TN_TC16_RUN=1 > $O/tc16_run_5b.log 2>&1
TC16_SHAPE=24,2,24 timeout 600 python tools/tc16_probe.py 1048576 - TN_TC16_RUN=1 > $O/tc16_run_3.log 2>&1
TN_TC16_RUN=1 timeout 300 python tools/tc_one.py 65536 f16 > $O/tc16_one_plain.log 2>&1 && \
  TN_TC16_RUN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tc16 -c 1 -o $O/ncu_tc16run python tools/tc_one.py 65536 f16 > $O/ncu_tc16run.log 2>&1
ncu -i $O/ncu_tc16run.ncu-rep --page raw --csv > $O/ncu_tc16run_raw.csv 2>/dev/null
python tools/ncu_hotspots.py $O/ncu_tc16run.ncu-rep 40 > $O/ncu_tc16run_hotspots.txt 2>&1
rm -f $O/ncu_tc16run.ncu-rep
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks > $O/bench_plain.json 2> $O/bench_plain.err; echo "bench plain rc=$?" >> $O/rc.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/launches_cfg5a_1M.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-peaks > $O/bench_ncu.json 2> $O/bench_ncu.err; echo "bench ncu rc=$?" >> $O/rc.txt
python tools/launch_summary.py $O/launches_cfg5a_1M.csv 40 > $O/launches_cfg5a_1M_summary.txt 2>&1
gzip -f $O/launches_cfg5a_1M.csv
echo done >> $O/rc.txt
