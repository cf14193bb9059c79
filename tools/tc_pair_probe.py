"""Compare the 1-CTA and the CTA-pair (cta_group::2) tcgen05 Gram kernels on a config-5a middle site: agreement and time."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
DEV = "cuda"
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ma, mb, mc = 38, 29, 38
g = torch.Generator(device=DEV).manual_seed(2)
fa = Factor(torch.randn((S, ma), device=DEV, generator=g), m=ma)
fb = Factor(torch.rand((S, mb), device=DEV, generator=g) * 2 - 1, m=mb)
fc = Factor(torch.randn((S, mc), device=DEV, generator=g), m=mc)
w = torch.rand((S,), device=DEV, generator=g) + 0.5
npair = lambda m: m * (m + 1) // 2
n = npair(ma) * npair(mb) * npair(mc)
flops = 2.0 * S * n


def run(no_pair, mode, reps=2):
    if no_pair:
        os.environ.pop("TN_TC_PAIR", None)
    else:
        os.environ["TN_TC_PAIR"] = "1"
    M = torch.empty(n, device=DEV)
    ops.gram(mode, fa, fb, fc, w, S, M=M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gram(mode, fa, fb, fc, w, S, M=M)
    e1.record(); torch.cuda.synchronize()
    return M, e0.elapsed_time(e1) / reps


for mode, nm, mult in ((ops.GRAM_TF32X3, "tf32x3", 3), (ops.GRAM_TF32, "tf32", 1)):
    M1, t1 = run(True, mode)
    M2, t2 = run(False, mode)
    d = float((M1 - M2).norm() / M1.norm())
    print(json.dumps({"mode": nm, "rows": S, "one_cta_ms": t1, "pair_ms": t2, "one_cta_tflops": flops * mult / t1 / 1e9,
                      "pair_tflops": flops * mult / t2 / 1e9, "rel_diff_pair_vs_one": d}), flush=True)
if S <= 8192:
    ref = ops.gram(ops.GRAM_FP64, fa, fb, fc, w, S)
    os.environ["TN_TC_PAIR"] = "1"
    got = ops.gram(ops.GRAM_TF32X3, fa, fb, fc, w, S)
    print(json.dumps({"pair_vs_fp64_rel": float((got - ref).norm() / ref.norm())}))
