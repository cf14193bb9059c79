#!/bin/bash
# round 2, GPU call 40 (1 GPU): bench.py after the config refactoring (both arms, short shape)
mkdir -p gpurun_out/r2c40; O=gpurun_out/r2c40
timeout 300 python bench.py --steps 1 --warmup 3 --rows 131072 --no-peaks > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" > $O/rc.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --ref-quick > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?" >> $O/rc.txt
echo done >> $O/rc.txt
