#!/bin/bash
# round 2, GPU call 24 (1 GPU): full GPU suite, default bench line (with CPU arm and e2e), ncu captures of the V-image Gram kernel
mkdir -p gpurun_out/r2c24; O=gpurun_out/r2c24
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" > $O/rc.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
export TN_TC_FLUSH_ROWS=16384
timeout 300 python tools/tc_one.py 131072 f16 > $O/tc_one_plain.log 2>&1 && \
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gram_tc16_vimg -s 1 -c 1 --csv --log-file $O/ncu_vimg_traffic.csv python tools/tc_one.py 131072 f16 > $O/ncu_vimg_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tc16_vimg -s 1 -c 1 -o $O/ncu_vimg python tools/tc_one.py 65536 f16 > $O/ncu_vimg.log 2>&1
ncu -i $O/ncu_vimg.ncu-rep --page raw --csv > $O/ncu_vimg_raw.csv 2>/dev/null
python tools/ncu_hotspots.py $O/ncu_vimg.ncu-rep 40 > $O/ncu_vimg_hotspots.txt 2>&1
rm -f $O/ncu_vimg.ncu-rep
echo done >> $O/rc.txt
