#!/bin/bash
# round 2, GPU call 6 (2 GPUs): sharded sweeps against the single-GPU sweep over NCCL, and the bench at N = 2
mkdir -p gpurun_out/r2c6; O=gpurun_out/r2c6
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py > $O/shard_check.jsonl 2> $O/shard_check.err; echo "shard_check rc=$?" > $O/rc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/rc.txt
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
