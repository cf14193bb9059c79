#!/bin/bash
# round 2, GPU call 28 (8 GPUs): the north-star target run (10 M rows, rank 38, degree 5) and the driver's default shape at N = 8 with the final kernels
mkdir -p gpurun_out/r2c28; O=gpurun_out/r2c28
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 400 $TR --master-port 29521 bench.py --gpus 8 --rows 1250000 --steps 2 --warmup 3 > $O/bench_T1_10M_n8.json 2> $O/bench_T1_10M_n8.err; echo "T1 rc=$?" > $O/rc.txt
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 3 > $O/bench_1M_n8.json 2> $O/bench_1M_n8.err; echo "default n8 rc=$?" >> $O/rc.txt
grep -h "NCCL INFO.*\(NVLS\|comm 0x.*nranks\)" $O/bench_T1_10M_n8.err | head -20 > $O/nccl_init_lines.txt
echo done >> $O/rc.txt
