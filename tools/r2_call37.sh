#!/bin/bash
# round 2, GPU call 37 (8 GPUs): the north-star target run once more with the final library
mkdir -p gpurun_out/r2c37; O=gpurun_out/r2c37
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 bench.py --gpus 8 --rows 1250000 --steps 2 --warmup 2 > $O/bench_T1_10M_n8.json 2> $O/bench_T1_10M_n8.err; echo "T1 rc=$?" > $O/rc.txt
echo done >> $O/rc.txt
