#!/bin/bash
# round 2, GPU call 9 (1 GPU): the warp-per-tile bulk-copy environment kernel (test, A/B timing, ncu), ncu --set full of the fp16 Gram kernel
mkdir -p gpurun_out/r2c9; O=gpurun_out/r2c9
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -rA -x -k env -p no:cacheprovider > $O/pytest_env.log 2>&1; echo "env tests rc=$?" > $O/rc.txt
timeout 300 python tools/tc_probe.py env > $O/env_probe_warp.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
TN_ENV_NO_WARP=1 timeout 300 python tools/tc_probe.py env > $O/env_probe_nowarp.log 2>&1
timeout 300 python tools/env_one.py > $O/env_one_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -c 6 -o $O/ncu_env python tools/env_one.py > $O/ncu_env.log 2>&1
timeout 300 python tools/tc_one.py 65536 f16 > $O/tc16_one_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tc16 -c 1 -o $O/ncu_tc16 python tools/tc_one.py 65536 f16 > $O/ncu_tc16.log 2>&1
ncu -i $O/ncu_tc16.ncu-rep --page raw --csv > $O/ncu_tc16_raw.csv 2>/dev/null
ncu -i $O/ncu_env.ncu-rep --page raw --csv > $O/ncu_env_raw.csv 2>/dev/null
echo done >> $O/rc.txt
