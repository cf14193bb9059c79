"""fp64 Gram (the bit-comparable mode): software-pipelined kernel against the three-barrier kernel (TN_GRAM_F64_NO_PIPE=1)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
g = torch.Generator(device="cuda").manual_seed(0)
npair = lambda m: m * (m + 1) // 2
for name, S, ma, mb, mc in (("cfg3_mid", 515345, 24, 2, 24), ("cfg5a_mid", 32768, 38, 29, 38), ("cfg5b_mid", 131072, 38, 6, 38), ("cfg2", 20640, 100, 9, 1), ("cfg1", 4177, 6, 9, 6)):
    Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
    w = torch.rand((S,), device="cuda", generator=g) + 0.5
    args = (ops.GRAM_FP64, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
    fl = 2.0 * S * npair(ma) * npair(mb) * npair(mc)
    outs = {}
    for var in ("pipe", "old"):
        if var == "old":
            os.environ["TN_GRAM_F64_NO_PIPE"] = "1"
        else:
            os.environ.pop("TN_GRAM_F64_NO_PIPE", None)
        M = ops.gram(*args); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.gram(*args, M=M)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        outs[var] = M.clone()
        print(json.dumps({"site": name, "rows": S, "kernel": var, "ms": ms, "tflops_fp64": fl / ms / 1e9}), flush=True)
    print(json.dumps({"site": name, "rel_diff": float((outs["pipe"] - outs["old"]).norm() / outs["old"].norm()), "same_bits": bool(torch.equal(outs["pipe"], outs["old"]))}), flush=True)
