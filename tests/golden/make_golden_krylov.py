"""Golden recordings of the reference's matrix-free sweeps (lanczos_swipe, scipy_swipe) -- build container only.

    python tests/golden/make_golden_krylov.py

lanczos_swipe draws its start vector with torch.randn_like (tensor/network.py:793); the draw is replaced here by a
recorded vector so the B200 path can be given the same x0.
"""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from scipy.sparse.linalg import cg, minres  # noqa: E402
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction, XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def data(seed, N, F, K=None):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    if K is None:
        y = torch.tensor(np.tanh(X[:, :1]) + 0.3 * X[:, 1:2] ** 2 + 0.05 * rng.normal(size=(N, 1)))
    else:
        y = torch.tensor(np.eye(K)[np.argmax(X @ rng.normal(size=(F, K)), axis=1)])
    return Xb, y


def record(kind, name, layer, X, y, loss_fn, **kw):
    tn = layer.tensor_network
    cores0 = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    ups = []
    x0s = []
    rng = np.random.default_rng(123)
    orig = torch.randn_like

    def fake_randn_like(t, *a, **k):
        v = torch.tensor(rng.normal(size=tuple(t.shape)))
        x0s.append(v.numpy().copy())
        return v

    def block_callback(NS, node):
        ups.append({"NS": NS, "k": tn.train_nodes.index(node), "after": [n.tensor.detach().numpy().copy() for n in tn.train_nodes]})

    losses = []
    if kind == "lanczos":
        torch.randn_like = fake_randn_like
        try:
            tn.lanczos_swipe(X, y, loss_fn, block_callback=block_callback, loss_callback=lambda l: losses.append(l), **kw)
        finally:
            torch.randn_like = orig
    else:
        solver = {"cg": cg, "minres": minres}[kw.pop("solver")]
        tn.scipy_swipe(X, y, loss_fn, solver, block_callback=block_callback, loss_callback=lambda l: losses.append(l), **kw)
    flat = {"x": X.numpy(), "y": y.numpy(), "n_cores": np.array(len(cores0)), "n_updates": np.array(len(ups)), "losses": np.array(losses)}
    for i, c in enumerate(cores0):
        flat[f"cores0_{i}"] = c
    for ui, u in enumerate(ups):
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"]])
        for i, c in enumerate(u["after"]):
            flat[f"u{ui}_after_{i}"] = c
        if kind == "lanczos":
            flat[f"u{ui}_x0"] = x0s[ui]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print(name, len(ups), "updates", "losses", losses[:3])


def main():
    X, y = data(1, 150, 3)
    layer = TensorTrainLayer(4, 3, 4, output_shape=1, constrict_bond=False, seed=9)
    record("lanczos", "krylov_lanczos_reg", layer, X, y, SquareBregFunction(), batch_size=50, num_swipes=2, lr=1.0, max_iter=6, tol=1e-12)
    X, y = data(2, 160, 3, K=3)
    layer = TensorTrainLayer(3, 3, 4, output_shape=2, constrict_bond=False, seed=4)
    record("lanczos", "krylov_lanczos_xe", layer, X, y, XEAutogradBregman(w=1.0), batch_size=80, num_swipes=1, lr=1.0, max_iter=5, tol=1e-12)
    for solver in ("cg", "minres"):
        X, y = data(3, 180, 3)
        layer = TensorTrainLayer(4, 3, 4, output_shape=1, constrict_bond=False, seed=2)
        record("scipy", f"krylov_scipy_{solver}", layer, X, y, SquareBregFunction(), solver=solver, batch_size=60, num_swipes=2, lr=1.0,
               max_iter=25, tol=1e-5)


if __name__ == "__main__":
    main()
