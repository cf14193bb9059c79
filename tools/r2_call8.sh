#!/bin/bash
# round 2, GPU call 8 (8 GPUs): the north-star target run (10 M rows, rank 38, degree 5), the driver's default shape at N = 8, config 3 sharded
mkdir -p gpurun_out/r2c8; O=gpurun_out/r2c8
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR --master-port 29521 bench.py --gpus 8 --rows 1250000 --steps 2 --warmup 3 > $O/bench_T1_10M_n8.json 2> $O/bench_T1_10M_n8.err; echo "T1 rc=$?" > $O/rc.txt
timeout 400 $TR --master-port 29522 bench.py --gpus 8 --steps 1 --warmup 1 > $O/bench_1M_n8.json 2> $O/bench_1M_n8.err; echo "default n8 rc=$?" >> $O/rc.txt
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --workload cfg3 --steps 2 --warmup 3 > $O/bench_cfg3_weak_n8.json 2> $O/bench_cfg3_weak_n8.err; echo "cfg3 weak rc=$?" >> $O/rc.txt
timeout 300 $TR --master-port 29524 bench.py --gpus 8 --workload cfg3 --rows 64419 --steps 2 --warmup 3 > $O/bench_cfg3_strong_n8.json 2> $O/bench_cfg3_strong_n8.err; echo "cfg3 strong rc=$?" >> $O/rc.txt
grep -h "NCCL INFO.*\(NVLS\|Connected\|Channel\|comm 0x.*nranks\|Using network\)" $O/bench_T1_10M_n8.err | head -40 > $O/nccl_init_lines.txt
echo done >> $O/rc.txt
