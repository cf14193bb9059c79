"""One launch of the tcgen05 Gram kernel on a config-5a middle site (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = {"tf32x3": ops.GRAM_TF32X3, "tf32": ops.GRAM_TF32, "fp64": ops.GRAM_FP64, "f16": ops.GRAM_F16}[sys.argv[2] if len(sys.argv) > 2 else "tf32x3"]
ma, mb, mc = (int(v) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else "38,29,38".split(",")))
g = torch.Generator(device="cuda").manual_seed(0)
Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
w = torch.full((S,), 2.0, device="cuda")
for _ in range(2):
    M = ops.gram(mode, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
torch.cuda.synchronize()
print("ok", float(M.sum()))
