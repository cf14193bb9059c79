#!/bin/bash
# round 2, GPU call 26 (1 GPU): env warp kernel with three tile buffers and direct stores
mkdir -p gpurun_out/r2c26; O=gpurun_out/r2c26
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k env -p no:cacheprovider > $O/pytest_env.log 2>&1; echo "env tests rc=$?" > $O/rc.txt
timeout 300 python tools/tc_probe.py env > $O/env_probe.log 2>&1; echo "probe rc=$?" >> $O/rc.txt
timeout 300 python tools/env_one.py > $O/env_one_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -c 6 -o $O/ncu_env python tools/env_one.py > $O/ncu_env.log 2>&1
python tools/ncu_summary.py $O/ncu_env.ncu-rep > $O/ncu_env_summary.txt 2>&1
rm -f $O/ncu_env.ncu-rep
echo done >> $O/rc.txt
