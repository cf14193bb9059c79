#!/bin/bash
# round 2, GPU call 5: the fp16-operand Gram kernel -- correctness first (in a child process with a timeout), then speed
mkdir -p gpurun_out/r2c5; O=gpurun_out/r2c5
timeout 300 python -m pytest tests/test_gpu_gram_tc.py -q -rA -x -p no:cacheprovider -k "f16 or matches_fp64" > $O/pytest_f16.log 2>&1; echo "f16 kernel rc=$?" > $O/rc.txt
if grep -q "f16 kernel rc=0" $O/rc.txt; then
  timeout 300 python tools/tc_one.py 131072 f16 > $O/tc_one_f16.log 2>&1
  timeout 600 python - > $O/tc_speed.log 2>&1 <<'PY'
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
ref = None
for name, mode in (("tf32", ops.GRAM_TF32), ("f16", ops.GRAM_F16), ("tf32x3", ops.GRAM_TF32X3)):
    args = (mode, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
    ops.gram(*args, M=M, flush_rows=8192); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        ops.gram(*args, M=M, flush_rows=8192)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(json.dumps({"mode": name, "rows": S, "ms": ms, "executed_tflops": fl * (3 if name == "tf32x3" else 1) / ms / 1e9, "useful_tflops": fl / ms / 1e9}), flush=True)
PY
  timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "suite rc=$?" >> $O/rc.txt
  timeout 600 python tools/cfg_margins.py > $O/cfg_margins.log 2>&1
  timeout 900 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --gram-mode f16 > $O/bench_1M_f16.json 2> $O/bench_1M_f16.err
  timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
  timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg4a --max-iter 50 > $O/bench_cfg4a.json 2> $O/bench_cfg4a.err
fi
echo done >> $O/rc.txt
