#!/bin/bash
# round 2, GPU call 35 (1 GPU): L2 policies of the V-image Gram kernel (flush reductions evict_last / evict_first, evict_last hint on the V copies): DRAM bytes and time
mkdir -p gpurun_out/r2c35; O=gpurun_out/r2c35
export TN_TC_FLUSH_ROWS=16384
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for v in "1 0" "0 0" "0 1" "1 1"; do
  set -- $v
  TN_TC16_RED_POLICY=$1 TN_TC16_STREAM_KEEP=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:gram_tc16_vimg -s 1 -c 1 --csv --log-file $O/traffic_red$1_keep$2.csv python tools/tc_one.py 131072 f16 > $O/ncu_red$1_keep$2.log 2>&1
done
timeout 300 python tools/tc16_probe.py 262144 - TN_TC16_RED_POLICY=0 TN_TC16_RED_POLICY=0,TN_TC16_STREAM_KEEP=1 TN_TC16_STREAM_KEEP=1 > $O/tc16_policy.log 2>&1; echo "probe rc=$?" > $O/rc.txt
echo done >> $O/rc.txt
