"""Loader for the fixtures written by tests/golden/make_golden.py."""
import ast
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith(("krylov_", "dmrg_", "cumsum_", "type1_", "conv_", "growing_", "linear_", "batch_", "grad_", "cfg1_", "cfg2_", "cfg3_", "cfg5b_", "cfg4b_", "cfg5a_", "cfg4a_", "wrappers")))


def load_krylov(name):
    """Fixtures of tests/golden/make_golden_krylov.py (lanczos_swipe / scipy_swipe recordings)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    nc = int(z["n_cores"])
    ups = []
    for ui in range(int(z["n_updates"])):
        NS, k = z[f"u{ui}_scal"]
        u = {"NS": int(NS), "k": int(k), "after": [z[f"u{ui}_after_{i}"] for i in range(nc)]}
        if f"u{ui}_x0" in z.files:
            u["x0"] = z[f"u{ui}_x0"]
        ups.append(u)
    return {"x": z["x"], "y": z["y"], "cores0": [z[f"cores0_{i}"] for i in range(nc)], "updates": ups, "losses": z["losses"]}


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["meta"]))
    nc = int(z["n_cores"])
    if bool(z["x_is_list"]):
        x = [z[f"x_{i}"] for i in range(sum(1 for k in z.files if k.startswith("x_") and k != "x_is_list"))]
    else:
        x = z["x"]
    ups = []
    for ui in range(int(z["n_updates"])):
        NS, k, eps, loss = z[f"u{ui}_scal"]
        u = {"NS": int(NS), "k": int(k), "eps": float(eps), "loss": float(loss),
             "A": z[f"u{ui}_A"], "b": z[f"u{ui}_b"], "step": z[f"u{ui}_step"],
             "before": [z[f"u{ui}_before_{i}"] for i in range(nc)],
             "after": [z[f"u{ui}_after_{i}"] for i in range(nc)]}
        for key in ("L", "R"):
            if f"u{ui}_{key}" in z.files:
                u[key] = z[f"u{ui}_{key}"]
        ups.append(u)
    return {"meta": meta, "x": x, "y": z["y"], "cores0": [z[f"cores0_{i}"] for i in range(nc)],
            "updates": ups, "pred": z["pred"], "ok": bool(z["ok"])}


def sweep_extras(meta):
    """The accumulating_swipe keywords a fixture sets beyond the common ones."""
    return {k: meta[k] for k in ("direction", "eps_per_node", "adaptive_step", "max_norm") if k in meta}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d
