#!/bin/bash
# round 2, GPU call 38 (1 GPU): ncu --set full of the factorisation's tensor-core update (after the bulk-copy ring) and of rhs_big_kernel (after the load-only prefetch)
mkdir -p gpurun_out/r2c38; O=gpurun_out/r2c38
timeout 300 python tools/chol_one.py 16384 mixed 1 > $O/chol_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_tc_kernel -s 16 -c 1 -o $O/ncu_syrk python tools/chol_one.py 16384 mixed 1 > $O/ncu_syrk.log 2>&1
python tools/ncu_summary.py $O/ncu_syrk.ncu-rep > $O/ncu_syrk_summary.txt 2>&1
python tools/ncu_hotspots.py $O/ncu_syrk.ncu-rep 15 > $O/ncu_syrk_hotspots.txt 2>&1
ncu -i $O/ncu_syrk.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k in ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_tc_wavefronts_mem_shared.sum','sm__cycles_elapsed.avg.per_second'):
    if k in h: print('  ', k, '=', r[h.index(k)])
" >> $O/ncu_syrk_summary.txt 2>&1
rm -f $O/ncu_syrk.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_big -c 1 -o $O/ncu_rhs_big python tools/rhs_probe.py > $O/ncu_rhs_big.log 2>&1
python tools/ncu_summary.py $O/ncu_rhs_big.ncu-rep > $O/ncu_rhs_big_summary.txt 2>&1
python tools/ncu_hotspots.py $O/ncu_rhs_big.ncu-rep 12 > $O/ncu_rhs_big_hotspots.txt 2>&1
rm -f $O/ncu_rhs_big.ncu-rep
echo done > $O/rc.txt
