#!/bin/bash
# round 2, GPU call 44 (1 GPU): 16-byte cp.async in the factored fp64 Gram kernel (against 8-byte copies and the unfactored kernel), GPU suite with
# the kernel's own tests, smoke
mkdir -p gpurun_out/r2c44; O=gpurun_out/r2c44
timeout 150 python tools/gram_f64_probe.py quick > $O/gram_f64_probe.log 2>&1; echo "probe rc=$?" > $O/rc.txt
timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
