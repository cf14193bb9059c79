"""fp16 Gram kernel variants on the config-5a middle site: time, executed TF/s, bits against the default variant.
usage: tc16_probe.py ROWS VAR=VAL[,VAR=VAL...] [VAR=VAL...]   (each argument = one variant = a set of TN_* switches; '-' = defaults)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
S = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
variants = sys.argv[2:] or ["-"]
shape = tuple(int(v) for v in os.environ.get("TC16_SHAPE", "38,29,38").split(","))
ma, mb, mc = shape
g = torch.Generator(device="cuda").manual_seed(0)
Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
w = torch.full((S,), 2.0, device="cuda")
npair = lambda m: m * (m + 1) // 2
fl = 2.0 * S * npair(ma) * npair(mb) * npair(mc)
args = (ops.GRAM_F16, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
base = None
for var in variants:
    sets = {} if var == "-" else dict(kv.split("=") for kv in var.split(","))
    for k, v in sets.items():
        os.environ[k] = v
    M = torch.empty(npair(ma) * npair(mb) * npair(mc), device="cuda")
    ops.gram(*args, M=M, flush_rows=8192); torch.cuda.synchronize()
    if base is None:
        base = M.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.gram(*args, M=M, flush_rows=8192)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"variant": var, "shape": shape, "rows": S, "ms": ms, "executed_tflops": fl / ms / 1e9, "same_bits_as_first": bool(torch.equal(M, base)),
                      "rel_diff_vs_first": float((M - base).norm() / base.norm())}), flush=True)
    for k in sets:
        os.environ.pop(k, None)
