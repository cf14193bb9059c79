"""Default vs opt-in variants of the tcgen05 Gram kernel on the config-5a middle site: time and agreement of M (GPU)."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
S, ma, mb, mc = int(sys.argv[2]), 38, 29, 38
g = torch.Generator(device="cuda").manual_seed(0)
Fa = torch.randn((S, ma), device="cuda", generator=g); Fb = torch.rand((S, mb), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
w = torch.full((S,), 2.0, device="cuda")
npair = lambda m: m * (m + 1) // 2
M = torch.empty(npair(ma) * npair(mb) * npair(mc), device="cuda")
fl = 2.0 * S * npair(ma) * npair(mb) * npair(mc) * 3
args = (ops.GRAM_TF32X3, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
ops.gram(*args, M=M); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    ops.gram(*args, M=M)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(json.dumps({"variant": sys.argv[3], "rows": S, "ms": ms, "issued_tflops": fl / ms / 1e9, "M_sum": float(M.sum()), "M_norm": float(M.norm())}))
'''
rows = sys.argv[1] if len(sys.argv) > 1 else "131072"
for name, env in (("default", {}), ("planar", {"TN_TC_RAW_PLANAR": "1"}), ("planar+vstage", {"TN_TC_RAW_PLANAR": "1", "TN_TC_V_PRESTAGE": "1"}),
                  ("pair", {"TN_TC_PAIR": "1"})):
    e = dict(os.environ, **env)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, ROOT, rows, name], capture_output=True, text=True, timeout=240, env=e)
        print(r.stdout.strip() or json.dumps({"variant": name, "error": r.stderr.strip().splitlines()[-3:]}), flush=True)
    except subprocess.TimeoutExpired:
        print(json.dumps({"variant": name, "error": "timeout"}), flush=True)
