"""One blocked Cholesky solve at size P (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lda = (P + 7) // 8 * 8
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.empty((P, lda), device="cuda")
A[:, :P] = 0.5 / P ** 0.5 * torch.randn((P, P), device="cuda", generator=g)
A[:, :P] = 0.5 * (A[:, :P] + A[:, :P].t())
A[:, :P].diagonal().add_(2.0)
r = torch.randn((P,), device="cuda", generator=g)
A0 = A.clone() if P <= 20000 else None
if A0 is not None:
    ops.cholesky_solve(A0.clone(), r.clone())        # warm-up: module load, attribute setup, allocator
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); info = ops.cholesky_solve(A, r); e1.record(); torch.cuda.synchronize()
print("ok", int(info.item()), e0.elapsed_time(e1), "ms", P ** 3 / 3 / e0.elapsed_time(e1) / 1e9, "TF/s")
