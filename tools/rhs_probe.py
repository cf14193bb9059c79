"""Right-hand side / J^T u pass on large cores: the register-blocked kernel against the GEMM-shaped one (TN_RHS_NO_BIG=1)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
g = torch.Generator(device="cuda").manual_seed(0)
for name, S, ma, mb, mc in (("cfg5a_mid", 1000000, 38, 29, 38), ("cfg5b_mid", 1000000, 38, 6, 38), ("r24_f12", 500000, 24, 12, 24)):
    Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
    w = torch.randn((S,), device="cuda", generator=g)
    fa, fb, fc = Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc)
    outs = {}
    for var in ("big", "gemm"):
        if var == "gemm":
            os.environ["TN_RHS_NO_BIG"] = "1"
        else:
            os.environ.pop("TN_RHS_NO_BIG", None)
        b = ops.rhs(fa, fb, fc, w, S); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.rhs(fa, fb, fc, w, S, b=b)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        outs[var] = b.clone()
        print(json.dumps({"site": name, "rows": S, "kernel": var, "ms": ms, "tflops_fp64": 2.0 * S * ma * mb * mc / ms / 1e9,
                          "GBs_factors": 8.0 * S * (ma + mb + mc + 1) / ms / 1e6}), flush=True)
    print(json.dumps({"site": name, "rel_diff": float((outs["big"] - outs["gemm"]).norm() / outs["gemm"].norm())}), flush=True)
