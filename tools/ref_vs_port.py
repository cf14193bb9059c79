"""Build container only (needs /root/reference): time the UNMODIFIED reference on the host cores next to the oracle port that
bench.py's cpu_baseline / --impl reference arm uses on the GPU box (where the Python reference cannot travel), on the same inputs.

    python tools/ref_vs_port.py > profiles/r1_reference_vs_port_cpu.json

Per case one site update = forward + get_A_b + solve_system + update on ONE minibatch:
  A  the reference verbatim (torch.einsum contracts J,J first: an S x P x P temporary, network.py:212);
  B  the reference with an `opt_einsum` stand-in on sys.path, as in its authors' environment ((J H) first);
  port  oracle/tn_oracle.py (numpy/BLAS), the code the bench times on the GPU box.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    "cfg1_mid_site": dict(n=3, r=6, f=9, site=1, rows=512, constrict=True),           # P = 324
    "cfg3_like_mid_site": dict(n=6, r=24, f=2, site=3, rows=256, constrict=False),    # P = 1152, as the middle of the 90-site train
    "cfg5b_like_site": dict(n=4, r=16, f=6, site=1, rows=256, constrict=False),       # P = 1536
}

CHILD = r'''
import sys, time, types, json
m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
if sys.argv[1] == "B":
    sys.path.insert(0, sys.argv[3])              # the opt_einsum stand-in must be importable before torch
sys.path.insert(0, "/root/reference")
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLayer
from tensor.bregman import SquareBregFunction
c = json.loads(sys.argv[2])
rng = np.random.default_rng(0)
X = torch.tensor(rng.uniform(-1, 1, size=(c["rows"], c["f"])))
y = torch.tensor(rng.normal(size=(c["rows"], 1)))
layer = TensorTrainLayer(c["n"], c["r"], c["f"], output_shape=1, constrict_bond=c["constrict"], seed=1)
tn = layer.tensor_network
node = tn.train_nodes[c["site"]]
best = None
for rep in range(3):
    t0 = time.perf_counter()
    tn.accumulating_swipe(X, y, SquareBregFunction(), node_order=[node], batch_size=-1, num_swipes=1, skip_second=True,
                          method="ridge_cholesky", eps=1.0)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print(json.dumps({"seconds": best, "P": int(node.tensor.numel()), "opt_einsum": bool(torch.backends.opt_einsum.is_available()),
                  "threads": torch.get_num_threads()}))
'''

STANDIN = '''"""Minimal stand-in for opt_einsum: torch only needs contract_path for its 3-operand einsum (SURVEY.md §8d, baseline B)."""
__version__ = "3.3.0"


def contract_path(*args, **kwargs):
    n = sum(1 for a in args[1:] if hasattr(a, "shape")) if isinstance(args[0], str) else len(args) // 2
    # three operands = the Gram einsum (J*, J, H): (J H) first, as opt_einsum would choose; otherwise left to right, pairwise
    path = [(1, 2), (0, 1)] if n == 3 else [(0, 1)] * max(n - 1, 1)
    return path, None
'''


def port_time(c):
    import numpy as np
    sys.path.insert(0, ROOT)
    import torch
    torch.set_default_dtype(torch.float64)
    import tensornetworksfork_b200 as tnb
    from oracle import tn_oracle as orc
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, size=(c["rows"], c["f"]))
    y = rng.normal(size=(c["rows"], 1))
    layer = tnb.TensorTrainLayer(c["n"], c["r"], c["f"], output_shape=1, constrict_bond=c["constrict"], seed=1)
    cores = [n.tensor.numpy().copy() for n in layer.tensor_network.train_nodes]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        orc.site_update([k.copy() for k in cores], X, y, c["site"], loss="square", batch_size=-1, method="ridge_cholesky", eps=1.0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def main():
    oe_dir = "/tmp/tn_opt_einsum_standin"
    os.makedirs(os.path.join(oe_dir, "opt_einsum"), exist_ok=True)
    open(os.path.join(oe_dir, "opt_einsum", "__init__.py"), "w").write(STANDIN)
    out = {"host": {"cores": os.cpu_count()}, "what": __doc__.strip().split("\n\n")[0], "cases": {}}
    for name, c in CASES.items():
        row = {"config": c}
        for arm in ("A", "B"):
            r = subprocess.run([sys.executable, "-c", CHILD, arm, json.dumps(c), oe_dir], capture_output=True, text=True, timeout=1800)
            if r.returncode != 0:
                row[arm] = {"error": r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "failed"}
            else:
                row[arm] = json.loads(r.stdout.strip().splitlines()[-1])
        row["port"] = {"seconds": port_time(c)}
        for arm in ("A", "B"):
            if "seconds" in row[arm]:
                row[f"{arm}_over_port"] = row[arm]["seconds"] / row["port"]["seconds"]
        out["cases"][name] = row
        print(name, json.dumps(row), file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
