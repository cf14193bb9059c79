"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (the reference lives at /root/reference there and
does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The reference stores no golden vectors of its own
(SURVEY.md §4, §8c), so these recordings of its live output are what pins both
the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py).

Each fixture records one ``accumulating_swipe`` run (reference: tensor/network.py:379)
teacher-forced: for every site update the cores *before* the update, the
accumulated ``A`` and ``b`` handed to ``solve_system`` (network.py:480), the
step it returned, the reported mean-of-batch-means loss (network.py:474,494)
and the cores *after* the update (and after the optional QR re-gauge).
"""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib")
p = types.ModuleType("matplotlib.pyplot")
m.pyplot = p
sys.modules["matplotlib"] = m
sys.modules["matplotlib.pyplot"] = p  # tensor/utils.py:2 imports it; plotting is unused
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.set_num_threads(4)

from tensor.layers import TensorTrainLayer, CPDLayer  # noqa: E402
from tensor.bregman import SquareBregFunction, XEAutogradBregman, AutogradLoss  # noqa: E402
from models.tnml import fbasis, polynomial_basis  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def canon_stack(stack, bond_label, out_labels):
    """Reference stack TensorNode -> (S, c, r) array, label addressed."""
    if stack is None:
        return None
    labels = list(stack.dim_labels)
    t = stack.tensor
    order = ["s"] + [l for l in out_labels if l != "s" and l in labels] + ([bond_label] if bond_label in labels else [])
    assert sorted(order) == sorted(labels), (order, labels)
    t = t.permute(*[labels.index(l) for l in order])
    S = t.shape[0]
    c = t.shape[1] if len(order) == 3 or (len(order) == 2 and bond_label not in labels) else 1
    return t.reshape(S, c, -1).detach().numpy().copy()


def record_swipe(layer, x, y, loss_fn, **kw):
    tn = layer.tensor_network
    rec = {"updates": []}
    cur = {}
    orig_solve = tn.solve_system
    orig_get_A_b = tn.get_A_b
    first_seen = set()

    def get_A_b(node, grad, hess, method=None):
        A, b = orig_get_A_b(node, grad, hess) if method is None else orig_get_A_b(node, grad, hess, method=method)
        if node.name not in first_seen and hasattr(tn, "get_stacks") and tn.left_stacks is not None:
            first_seen.add(node.name)
            ls, rs = tn.get_stacks(node)
            cur["L"] = canon_stack(ls, node.left_labels[0] if node.left_labels else None, tn.output_labels)
            cur["R"] = canon_stack(rs, node.right_labels[0] if node.right_labels else None, tn.output_labels)
        return A, b

    def solve_system(node, A, b, method="exact", eps=0.0):
        cur["cores_before"] = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
        cur["A"] = A.detach().numpy().copy()
        cur["b"] = b.detach().numpy().copy()
        cur["eps"] = float(eps)
        cur["method"] = method
        step = orig_solve(node, A, b, method=method, eps=eps)
        cur["step"] = step.detach().numpy().copy()
        return step

    def block_callback(NS, node):
        u = dict(cur)
        cur.clear()
        u["NS"] = NS
        u["k"] = tn.train_nodes.index(node)
        u["cores_after"] = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
        rec["updates"].append(u)

    def loss_callback(NS, node, loss):
        cur["loss"] = float(loss)

    tn.solve_system = solve_system
    tn.get_A_b = get_A_b
    rec["cores0"] = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    ok = tn.accumulating_swipe(x, y, loss_fn, block_callback=block_callback, loss_callback=loss_callback, **kw)
    rec["ok"] = bool(ok)
    bs = kw.get("batch_size", -1)
    rec["pred"] = tn.forward_batch(x, bs if bs > 0 else -1).detach().numpy().copy()
    return rec


def save(name, rec, x, y, meta):
    flat = {"meta": np.array(repr(meta)), "ok": np.array(rec["ok"]), "pred": rec["pred"], "y": y.numpy(),
            "n_updates": np.array(len(rec["updates"])), "n_cores": np.array(len(rec["cores0"]))}
    if isinstance(x, (list, tuple)):
        flat["x_is_list"] = np.array(True)
        for i, t in enumerate(x):
            flat[f"x_{i}"] = t.numpy()
    else:
        flat["x_is_list"] = np.array(False)
        flat["x"] = x.numpy()
    for i, c in enumerate(rec["cores0"]):
        flat[f"cores0_{i}"] = c
    for ui, u in enumerate(rec["updates"]):
        for key in ("A", "b", "step", "L", "R"):
            if u.get(key) is not None:
                flat[f"u{ui}_{key}"] = u[key]
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"], u["eps"], u["loss"]])
        for i, c in enumerate(u["cores_before"]):
            flat[f"u{ui}_before_{i}"] = c
        for i, c in enumerate(u["cores_after"]):
            flat[f"u{ui}_after_{i}"] = c
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"{name}: {len(rec['updates'])} updates, ok={rec['ok']}, {os.path.getsize(path) / 1024:.0f} KiB")


def teacher(X, seed, C=1):
    """Smooth synthetic target: random low-degree polynomial features + noise."""
    rng = np.random.default_rng(seed)
    W1 = rng.normal(size=(X.shape[1], C)) / np.sqrt(X.shape[1])
    W2 = rng.normal(size=(X.shape[1], C)) / np.sqrt(X.shape[1])
    return np.tanh(X @ W1) + 0.5 * (X @ W2) ** 2 + 0.05 * rng.normal(size=(X.shape[0], C))


def main():
    rng = np.random.default_rng(0)

    # 1. config-1 shape scaled down: TT poly-mode, perturb init, ridge_cholesky, minibatches
    N, F = 300, 4
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(teacher(X, 1))
    layer = TensorTrainLayer(3, 4, F + 1, output_shape=1, constrict_bond=True, perturb=True, seed=42)
    kw = dict(batch_size=128, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)
    rec = record_swipe(layer, Xb, y, SquareBregFunction(), **kw)
    save("tt_poly_reg", rec, Xb, y, dict(kind="tt", n=3, r=4, f=F + 1, C=1, loss="square", perturb=True, **kw))

    # 2. 5-site poly-mode TT, random init, unconstricted bonds (config-5a shape scaled down), full batch
    N, F = 257, 3
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(teacher(X, 2))
    layer = TensorTrainLayer(5, 3, F + 1, output_shape=1, constrict_bond=False, perturb=False, seed=7)
    kw = dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=[0.5, 0.1])
    rec = record_swipe(layer, Xb, y, SquareBregFunction(), **kw)
    save("tt_poly5_full", rec, Xb, y, dict(kind="tt", n=5, r=3, f=F + 1, C=1, loss="square", perturb=False, **kw))

    # 3. TNML sin-cos, orthonormalize=True (config-3 shape scaled down)
    N, F = 200, 6
    X = rng.uniform(-1, 1, size=(N, F))
    Xl = fbasis(torch.tensor(X))
    y = torch.tensor(teacher(X, 3))
    layer = TensorTrainLayer(F, 4, 2, output_shape=1, constrict_bond=True, perturb=False, seed=42)
    layer.tensor_network.orthonormalize_left()
    kw = dict(batch_size=64, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5, orthonormalize=True)
    rec = record_swipe(layer, Xl, y, SquareBregFunction(), **kw)
    save("tnml_sincos_qr", rec, Xl, y, dict(kind="tt", n=F, r=4, f=2, C=1, loss="square", basis="sin-cos", raw_x=X.tolist(), **kw))

    # 4. TNML polynomial basis degree 2, classifier with XE loss, class leg C=3 on site 1 (config-4a scaled down)
    N, F, K = 240, 5, 4
    X = rng.uniform(-1, 1, size=(N, F))
    Xl = polynomial_basis(torch.tensor(X), degree=2)
    labels = np.argmax(X @ rng.normal(size=(F, K)), axis=1)
    y = torch.tensor(np.eye(K)[labels])
    layer = TensorTrainLayer(F, 3, 3, output_shape=K - 1, constrict_bond=True, perturb=False, seed=5)
    kw = dict(batch_size=100, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.5)
    rec = record_swipe(layer, Xl, y, XEAutogradBregman(w=1.0), **kw)
    save("tnml_poly_xe", rec, Xl, y, dict(kind="tt", n=F, r=3, f=3, C=K - 1, loss="xe", w=1.0, basis="polynomial", degree=2, **kw))

    # 5. multi-output regression with the square loss quirk (H broadcast to all-ones*2) and with AutogradLoss(MSE)
    N, F, C = 150, 3, 2
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(teacher(X, 4, C))
    for nm, lf, lname in (("tt_multi_square", SquareBregFunction(), "square"), ("tt_multi_mse", AutogradLoss(), "mse")):
        layer = TensorTrainLayer(3, 3, F + 1, output_shape=C, constrict_bond=False, perturb=False, seed=11)
        kw = dict(batch_size=64, num_swipes=1, lr=1.0, method="ridge_exact", eps=0.3)
        rec = record_swipe(layer, Xb, y, lf, **kw)
        save(nm, rec, Xb, y, dict(kind="tt", n=3, r=3, f=F + 1, C=C, loss=lname, **kw))

    # 6. CPD rank 6, 4 factors (config-2 shape scaled down), AutogradLoss as default_CPD_house.py uses
    N, F = 220, 3
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(teacher(X, 6))
    layer = CPDLayer(4, 6, F + 1, output_shape=(1,), seed=42)
    kw = dict(batch_size=100, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.2)
    rec = record_swipe(layer, Xb, y, SquareBregFunction(), **kw)
    save("cpd_reg", rec, Xb, y, dict(kind="cpd", n=4, r=6, f=F + 1, C=1, loss="square", **kw))

    # 7. methods: exact / cholesky(with tiny ridge via eps ignored) / gradient-free variants on one small TT
    N, F = 120, 2
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(teacher(X, 8))
    layer = TensorTrainLayer(2, 2, F + 1, output_shape=1, constrict_bond=False, perturb=False, seed=3)
    kw = dict(batch_size=-1, num_swipes=1, lr=0.5, method="exact", eps=0.0, skip_second=True)
    rec = record_swipe(layer, Xb, y, SquareBregFunction(), **kw)
    save("tt_exact_lr", rec, Xb, y, dict(kind="tt", n=2, r=2, f=F + 1, C=1, loss="square", **kw))


if __name__ == "__main__":
    main()
