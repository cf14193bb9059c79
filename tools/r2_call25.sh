#!/bin/bash
# round 2, GPU call 25 (1 GPU): software-pipelined fp64 Gram kernel (tests, probe); shared-memory wavefronts per source line of the V-image Gram kernel
mkdir -p gpurun_out/r2c25; O=gpurun_out/r2c25
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -q -x -p no:cacheprovider > $O/pytest_fp64.log 2>&1; echo "tests rc=$?" > $O/rc.txt
timeout 300 python tools/gram_f64_probe.py > $O/gram_f64_probe.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
export TN_TC_FLUSH_ROWS=16384
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tc16_vimg -s 1 -c 1 -o $O/ncu_vimg python tools/tc_one.py 65536 f16 > $O/ncu_vimg.log 2>&1
python tools/ncu_hotspots.py $O/ncu_vimg.ncu-rep 12 > $O/ncu_vimg_hotspots.txt 2>&1
ncu -i $O/ncu_vimg.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | head -3 > $O/ncu_source_header.txt
rm -f $O/ncu_vimg.ncu-rep
echo done >> $O/rc.txt
