#!/bin/bash
# round 2, GPU call 10 (1 GPU): conflict-free core slab + padded per-row bulk copies in the env kernels; the fp16 CTA-pair Gram kernel
mkdir -p gpurun_out/r2c10; O=gpurun_out/r2c10
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -rA -x -k env -p no:cacheprovider > $O/pytest_env.log 2>&1; echo "env tests rc=$?" > $O/rc.txt
timeout 300 python tools/tc_probe.py env > $O/env_probe_warp.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
TN_ENV_NO_WARP=1 timeout 300 python tools/tc_probe.py env > $O/env_probe_nowarp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -k "f16" -p no:cacheprovider > $O/pytest_f16.log 2>&1; echo "f16 tests rc=$?" >> $O/rc.txt
timeout 600 python tools/tc16_probe.py 262144 - TN_TC16_PAIR=1 > $O/tc16_pair.log 2>&1; echo "pair probe rc=$?" >> $O/rc.txt
timeout 300 python tools/env_one.py > $O/env_one_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -c 6 -o $O/ncu_env python tools/env_one.py > $O/ncu_env.log 2>&1
python tools/ncu_summary.py $O/ncu_env.ncu-rep > $O/ncu_env_summary.txt 2>&1
rm -f $O/ncu_env.ncu-rep
echo done >> $O/rc.txt
