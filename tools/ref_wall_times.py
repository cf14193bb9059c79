"""Build container only (needs /root/reference): wall time of the UNMODIFIED reference's own sweeps at the BASELINE configurations,
on the host cores, around `accumulating_swipe` only.  Path B = the reference with `opt_einsum` importable
(a stand-in on sys.path, tools/ref_vs_port.py), its authors' setup; the verbatim path A (no opt_einsum) is timed where it is
feasible.  Data and models are those of tests/golden/make_golden_cfg{1,2,3,5b}.py.

    python tools/ref_wall_times.py > profiles/r1_reference_wall_times_cpu.json
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, time, types, json, os
arm, which = sys.argv[1], sys.argv[2]
root = sys.argv[3]
sys.path.insert(0, os.path.join(root, "tools"))
import ref_vs_port
if arm == "B":
    d = "/tmp/tn_opt_einsum_standin"
    os.makedirs(os.path.join(d, "opt_einsum"), exist_ok=True)
    open(os.path.join(d, "opt_einsum", "__init__.py"), "w").write(ref_vs_port.STANDIN)
    sys.path.insert(0, d)
m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(root, "tests"))
sys.path.insert(0, root)
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLayer, CPDLayer
from tensor.bregman import SquareBregFunction
n_up = [0]
cb = lambda NS, nd, l: n_up.__setitem__(0, n_up[0] + 1)
if which == "cfg1":
    import cfg1_case as c
    X, y = c.data()
    layer = TensorTrainLayer(3, c.R, c.F + 1, output_shape=1, constrict_bond=True, perturb=True, seed=42)
    x = torch.tensor(X); rows = X.shape[0]
    kw = dict(batch_size=512, lr=1.0, eps=c.EPSS, method="ridge_cholesky", num_swipes=c.NUM_SWIPES)
elif which == "cfg2":
    import cfg2_case as c
    X, y = c.data()
    layer = CPDLayer(c.FACTORS, c.RANK, c.F + 1, output_shape=(1,), seed=42)
    x = torch.tensor(X); rows = X.shape[0]
    kw = dict(batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5, method="ridge_cholesky", num_swipes=c.NUM_SWIPES)
elif which == "cfg3":
    import cfg3_case as c
    X, y = c.data()
    x = [torch.tensor(np.stack([np.cos(0.5 * np.pi * X[:, j]), np.sin(0.5 * np.pi * X[:, j])], 1)) for j in range(c.F)]
    layer = TensorTrainLayer(c.F, c.R, 2, output_shape=1, constrict_bond=True, seed=42)
    layer.tensor_network.orthonormalize_left(); rows = X.shape[0]
    kw = dict(batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5, orthonormalize=True, method="ridge_cholesky", num_swipes=1)
else:
    import cfg5b_case as c
    X, y = c.data()
    x = [torch.tensor(np.stack([X[:, j] ** d for d in range(c.DEG + 1)], 1)) for j in range(c.F)]
    layer = TensorTrainLayer(c.F, c.R, c.DEG + 1, output_shape=1, constrict_bond=True, seed=42)
    layer.tensor_network.orthonormalize_left(); rows = X.shape[0]
    kw = dict(batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5, orthonormalize=True, method="ridge_cholesky", num_swipes=1)
t0 = time.perf_counter()
ok = layer.tensor_network.accumulating_swipe(x, torch.tensor(y), SquareBregFunction(), loss_callback=cb, **kw)
dt = time.perf_counter() - t0
print(json.dumps({"seconds": dt, "site_updates": n_up[0], "rows": rows, "site_updates_per_s": n_up[0] / dt,
                  "sample_site_updates_per_s": n_up[0] * rows / dt, "ok": bool(ok), "opt_einsum": bool(torch.backends.opt_einsum.is_available()),
                  "threads": torch.get_num_threads()}))
'''

RUNS = [("cfg1", "A"), ("cfg1", "B"), ("cfg2", "B"), ("cfg3", "B"), ("cfg5b", "B")]
WHAT = {"cfg1": "config 1 at full size: 4177 x 9, 3 cores, rank 6, 4 sweeps (default_train.py call)",
        "cfg2": "config 2 at full size: CPD rank 100, 20640 x 9, 5 factors, 2 sweeps",
        "cfg3": "config 3 chain: 90 sites, sin-cos, rank 24, QR, one sweep on a 4096-row subsample (of 515k)",
        "cfg5b": "config 5b chain: 28 sites, polynomial degree 5, rank 38, QR, one sweep on a 2048-row subsample (of 1M-10M)"}


def main():
    out = {"host": {"cores": os.cpu_count()}, "what": __doc__.strip().split("\n\n")[0], "runs": []}
    for which, arm in RUNS:
        r = subprocess.run([sys.executable, "-c", CHILD, arm, which, ROOT], capture_output=True, text=True, timeout=3600)
        row = {"config": which, "path": arm, "workload": WHAT[which]}
        if r.returncode == 0:
            row.update(json.loads(r.stdout.strip().splitlines()[-1]))
        else:
            row["error"] = (r.stderr.strip().splitlines() or ["failed"])[-1]
        out["runs"].append(row)
        print(json.dumps(row), file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
