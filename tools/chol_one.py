"""One blocked Cholesky solve at size P (for timing / ncu).  usage: chol_one.py P [fp64|mixed] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = sys.argv[2] if len(sys.argv) > 2 else "fp64"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lda = (P + 7) // 8 * 8
g = torch.Generator(device="cuda").manual_seed(0)


def make():
    A = torch.empty((P, lda), device="cuda")
    A[:, :P] = 0.5 / P ** 0.5 * torch.randn((P, P), device="cuda", generator=g)
    A[:, :P] = 0.5 * (A[:, :P] + A[:, :P].t())
    A[:, :P].diagonal().add_(2.0)
    return A


def run(A, r):
    if mode == "mixed":
        return ops.cholesky_solve_mixed(A, r, rtol=1e-11)
    return ops.cholesky_solve(A, r), None


# warm-up at a small size: module load, attribute setup, allocator
Pw = 2048
Aw = torch.eye(Pw, device="cuda") * 2.0
run(Aw, torch.ones(Pw, device="cuda"))
torch.cuda.synchronize()
for rep in range(reps):
    A = make()
    r = torch.randn((P,), device="cuda", generator=g)
    check = P <= 24000
    if check:
        A0, r0 = A[:, :P].clone(), r.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); info, stats = run(A, r); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    msg = f"P {P} {mode}: info {int(info.item())} {ms:.1f} ms  {P ** 3 / 3 / ms / 1e9:.1f} TF/s (P^3/3)"
    if stats is not None:
        msg += f"  rel.resid {stats[0].item():.2e} after {int(stats[1].item())} refinement iterations"
    if check:
        msg += f"  true resid {float(torch.norm(A0 @ r - r0) / torch.norm(r0)):.2e}"
    print(msg, flush=True)
    del A
