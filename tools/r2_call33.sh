#!/bin/bash
# round 2, GPU call 33 (1 GPU): block-inverse substitutions: the solve tests, then the whole GPU suite
mkdir -p gpurun_out/r2c33; O=gpurun_out/r2c33
timeout 900 python -m pytest tests/test_gpu_solve_mixed.py -q -p no:cacheprovider > $O/pytest_solve.log 2>&1; echo "solve tests rc=$?" > $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" >> $O/rc.txt
timeout 300 python tools/chol_one.py 41876 mixed 2 > $O/chol_41876_mixed.log 2>&1
echo done >> $O/rc.txt
