import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
S, ma, mb, mc = 131072, 38, 29, 38
g = torch.Generator(device="cuda").manual_seed(0)
Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
w = torch.full((S,), 2.0, device="cuda")
npair = lambda m: m * (m + 1) // 2
M = torch.empty(npair(ma) * npair(mb) * npair(mc), device="cuda")
fl = 2.0 * S * npair(ma) * npair(mb) * npair(mc) * 3
for fr in (512, 1024, 2048, 4096, 8192):
    os.environ["TN_TC_FLUSH_ROWS"] = str(fr)
    ops.gram(ops.GRAM_TF32X3, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S, M=M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        ops.gram(ops.GRAM_TF32X3, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S, M=M)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(json.dumps({"flush_rows": fr, "ms": ms, "issued_tflops": fl / ms / 1e9}))
