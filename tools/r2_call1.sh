#!/bin/bash
# round 2, GPU call 1: state of the suite with the xfail markers gone, margins of the tf32x3 twins, the kernel variants written blind,
# the merged-loop bench on a short shape, the real reference arm on the box's host, and first ncu captures of the env / fp64 Gram kernels
mkdir -p gpurun_out/r2c1; O=gpurun_out/r2c1
(nproc; free -g; nvidia-smi -L; python -c "import psutil;print(psutil.virtual_memory())") > $O/host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/host.txt
timeout 600 python tools/cfg_margins.py > $O/cfg_margins.log 2>&1
timeout 900 python tools/tc_variants.py 131072 > $O/tc_variants.log 2>&1
TN_TEST_EXPERIMENTAL_KERNELS=1 timeout 300 python -m pytest tests/test_zz_gpu_late.py -q -rA -k planar -p no:cacheprovider > $O/pytest_planar.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --rows 131072 > $O/bench_131k.json 2> $O/bench_131k.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 python tools/tc_probe.py env > $O/env_probe.log 2>&1
timeout 300 python tools/env_one.py > $O/env_one_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -c 6 -o $O/ncu_env python tools/env_one.py > $O/ncu_env.log 2>&1
timeout 300 python tools/tc_one.py 65536 fp64 24,2,24 > $O/f64_one_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:kr3_f64 -c 2 -o $O/ncu_kr3f64 python tools/tc_one.py 65536 fp64 24,2,24 > $O/ncu_kr3f64.log 2>&1
echo done >> $O/host.txt
