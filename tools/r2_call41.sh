#!/bin/bash
# round 2, GPU call 41 (1 GPU): factored-operand fp64 Gram kernel -- correctness on awkward shapes, timing against the unfactored kernel, GPU suite
mkdir -p gpurun_out/r2c41; O=gpurun_out/r2c41
timeout 150 python tools/gram_f64_probe.py > $O/gram_f64_probe.log 2>&1; echo "probe rc=$?" > $O/rc.txt
timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
