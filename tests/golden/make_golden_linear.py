"""Golden recordings of the reference's TensorTrainLinearLayer (tensor/layers.py:308-343) under accumulating_swipe and
lanczos_swipe -- build container only.    python tests/golden/make_golden_linear.py"""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLinearLayer  # noqa: E402
from tensor.bregman import SquareBregFunction, XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def data(seed, N, F, K=None):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    if K is None:
        y = torch.tensor(np.tanh(X[:, :1] - X[:, 1:2]) + 0.3 * X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1)))
    else:
        y = torch.tensor(np.eye(K)[np.argmax(X @ rng.normal(size=(F, K)), axis=1)])
    return Xb, y


def record(name, layer, X, y, loss_fn, kind="dense", **kw):
    tn = layer.tensor_network
    flat = {"x": X.numpy(), "y": y.numpy(), "names": np.array([n.name for n in tn.train_nodes])}
    cores0 = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    flat["pred0"] = tn.forward(X, to_tensor=True).detach().numpy().copy()
    tn.reset_stacks()
    ups, losses, x0s = [], [], []

    def block_callback(NS, node):
        ups.append({"NS": NS, "k": tn.train_nodes.index(node), "after": [n.tensor.detach().numpy().copy() for n in tn.train_nodes]})

    if kind == "dense":
        ok = tn.accumulating_swipe(X, y, loss_fn, block_callback=block_callback, loss_callback=lambda NS, node, l: losses.append(l), **kw)
    else:
        rng = np.random.default_rng(77)
        orig = torch.randn_like

        def fake(t, *a, **k):
            v = torch.tensor(rng.normal(size=tuple(t.shape)))
            x0s.append(v.numpy().copy())
            return v
        torch.randn_like = fake
        try:
            ok = tn.lanczos_swipe(X, y, loss_fn, block_callback=block_callback, loss_callback=lambda l: losses.append(l), **kw)
        finally:
            torch.randn_like = orig
    tn.reset_stacks()
    flat["pred1"] = tn.forward(X, to_tensor=True).detach().numpy().copy()
    flat.update(ok=np.array(bool(ok)), n_cores=np.array(len(cores0)), n_updates=np.array(len(ups)), losses=np.array(losses))
    for i, c in enumerate(cores0):
        flat[f"cores0_{i}"] = c
    for ui, u in enumerate(ups):
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"]])
        for i, c in enumerate(u["after"]):
            flat[f"u{ui}_after_{i}"] = c
        if x0s:
            flat[f"u{ui}_x0"] = x0s[ui]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print(name, list(flat["names"]), len(ups), "updates", "losses", losses[:3], "->", losses[-1])


def main():
    X, y = data(1, 260, 5)
    layer = TensorTrainLinearLayer(3, 3, 6, 4, output_shape=1, constrict_bond=False, seed=5)
    record("linear_tt_reg", layer, X, y, SquareBregFunction(), batch_size=100, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0,
           eps_decay=0.5)
    X, y = data(2, 240, 4, K=3)
    layer = TensorTrainLinearLayer(3, 3, 5, 3, output_shape=2, constrict_bond=False, seed=6)
    record("linear_tt_xe", layer, X, y, XEAutogradBregman(w=1.0), batch_size=-1, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.5)
    X, y = data(3, 200, 4)
    layer = TensorTrainLinearLayer(3, 3, 5, 3, output_shape=1, constrict_bond=False, seed=7)
    record("linear_tt_lanczos", layer, X, y, SquareBregFunction(), kind="lanczos", batch_size=80, num_swipes=2, lr=1.0, max_iter=5, tol=1e-12)


if __name__ == "__main__":
    main()
