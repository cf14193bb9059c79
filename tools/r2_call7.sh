#!/bin/bash
# round 2, GPU call 7 (1 GPU): the one-launch matvec, 64-sample stages of the fp16 Gram kernel, small workloads again
mkdir -p gpurun_out/r2c7; O=gpurun_out/r2c7
timeout 600 python -m pytest tests/test_gpu_krylov_drivers.py tests/test_gpu_krylov.py tests/test_gpu_kernels.py -q -rA -x -p no:cacheprovider > $O/pytest_matvec.log 2>&1; echo "matvec tests rc=$?" > $O/rc.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" >> $O/rc.txt
for kc in 32 64; do
TN_TC16_KC=$kc timeout 300 python - >> $O/tc16_kc.log 2>&1 <<'PY'
import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
S, ma, mb, mc = 262144, 38, 29, 38
g = torch.Generator(device="cuda").manual_seed(0)
Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
w = torch.full((S,), 2.0, device="cuda")
npair = lambda m: m * (m + 1) // 2
M = torch.empty(npair(ma) * npair(mb) * npair(mc), device="cuda")
fl = 2.0 * S * npair(ma) * npair(mb) * npair(mc)
args = (ops.GRAM_F16, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
ref = ops.gram(ops.GRAM_TF32X3, *args[1:], flush_rows=2048)
ops.gram(*args, M=M, flush_rows=8192); torch.cuda.synchronize()
err = float((M - ref).norm() / ref.norm())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.gram(*args, M=M, flush_rows=8192)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"kc": os.environ.get("TN_TC16_KC"), "rows": S, "ms": ms, "executed_tflops": fl / ms / 1e9, "rel_err_vs_3xtf32": err}), flush=True)
PY
done
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg4a --max-iter 50 > $O/bench_cfg4a.json 2> $O/bench_cfg4a.err
TN_NO_FUSED_MATVEC=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg4a --max-iter 50 > $O/bench_cfg4a_twopass.json 2> $O/bench_cfg4a_twopass.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg5b > $O/bench_cfg5b.json 2> $O/bench_cfg5b.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg2 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg1 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
echo done >> $O/rc.txt
