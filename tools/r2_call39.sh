#!/bin/bash
# round 2, GPU call 39 (1 GPU): per-launch time and tensor-pipe activity of the factorisation's tensor-core updates (P = 16384, mixed solve)
mkdir -p gpurun_out/r2c39; O=gpurun_out/r2c39
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:syrk_tc_kernel -c 80 --csv --log-file $O/syrk_launches.csv python tools/chol_one.py 16384 mixed 1 > $O/ncu.log 2>&1
echo done > $O/rc.txt
