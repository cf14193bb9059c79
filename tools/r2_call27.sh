#!/bin/bash
# round 2, GPU call 27 (2 GPUs): sharded sweeps against the single-GPU sweep over NCCL with the final kernels; bench at N = 2 in the driver's form
mkdir -p gpurun_out/r2c27; O=gpurun_out/r2c27
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py > $O/shard_check.jsonl 2> $O/shard_check.err; echo "shard_check rc=$?" > $O/rc.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2.time; echo "bench n2 rc=$?" >> $O/rc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
