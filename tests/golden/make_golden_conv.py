"""Golden recordings of the reference's patch/pixel conv-TT layer (TensorConvolutionTrainLayer, tensor/layers.py:791-890)
under its matrix-free sweeps -- build container only.

    python tests/golden/make_golden_conv.py

Same protocol as make_golden_krylov.py: lanczos_swipe's random start vector (tensor/network.py:793) is replaced by a
recorded draw so the B200 path can be given the same x0.
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
from tensor.layers import TensorConvolutionTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction, XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def data(seed, N, Q, T, K=None, C=1):
    """Patches x pixels input with the bias patch / bias pixel of image_convolution_CG_MNIST.py:29-32."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(N, Q, T))
    X[:, -1, :] = 0.0
    X[:, :, -1] = 0.0
    X[:, -1, -1] = 1.0
    feat = X[:, :-1, :-1].reshape(N, -1)
    if K is None:
        W = rng.normal(size=(feat.shape[1], C)) / feat.shape[1] ** 0.5
        y = np.tanh(feat @ W) + 0.05 * rng.normal(size=(N, C))
    else:
        y = np.eye(K)[np.argmax(feat @ rng.normal(size=(feat.shape[1], K)), axis=1)]
    return torch.tensor(X), torch.tensor(y)


def record(kind, name, layer, X, y, loss_fn, **kw):
    tn = layer.tensor_network
    cores0 = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    names = [n.name for n in tn.train_nodes]
    pred0 = tn.forward(X, to_tensor=True).detach().numpy().copy()
    tn.reset_stacks()
    ups, x0s, losses = [], [], []
    rng = np.random.default_rng(321)
    orig = torch.randn_like

    def fake_randn_like(t, *a, **k):
        v = torch.tensor(rng.normal(size=tuple(t.shape)))
        x0s.append(v.numpy().copy())
        return v

    def block_callback(NS, node):
        ups.append({"NS": NS, "k": tn.train_nodes.index(node), "after": [n.tensor.detach().numpy().copy() for n in tn.train_nodes]})

    if kind == "lanczos":
        torch.randn_like = fake_randn_like
        try:
            tn.lanczos_swipe(X, y, loss_fn, block_callback=block_callback, loss_callback=lambda l: losses.append(l), **kw)
        finally:
            torch.randn_like = orig
    else:
        solver = {"cg": cg, "minres": minres}[kw.pop("solver")]
        tn.scipy_swipe(X, y, loss_fn, solver, block_callback=block_callback, loss_callback=lambda l: losses.append(l), **kw)
    tn.reset_stacks()
    pred1 = tn.forward(X, to_tensor=True).detach().numpy().copy()
    flat = {"x": X.numpy(), "y": y.numpy(), "n_cores": np.array(len(cores0)), "n_updates": np.array(len(ups)),
            "losses": np.array(losses), "names": np.array(names), "pred0": pred0, "pred1": pred1}
    for i, c in enumerate(cores0):
        flat[f"cores0_{i}"] = c
    for ui, u in enumerate(ups):
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"]])
        for i, c in enumerate(u["after"]):
            flat[f"u{ui}_after_{i}"] = c
        if kind == "lanczos":
            flat[f"u{ui}_x0"] = x0s[ui]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print(name, names, len(ups), "updates", "losses", losses[:3], "pred0", pred0.shape)


def main():
    torch.manual_seed(5)
    X, y = data(1, 120, 5, 4, K=3)
    layer = TensorConvolutionTrainLayer(num_carriages=3, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2)
    record("lanczos", "conv_lanczos_xe", layer, X, y, XEAutogradBregman(w=1.0), batch_size=50, num_swipes=2, lr=1.0, max_iter=6, tol=1e-12)
    torch.manual_seed(6)
    X, y = data(2, 150, 6, 5, C=1)
    layer = TensorConvolutionTrainLayer(num_carriages=4, bond_dim=3, num_patches=6, patch_pixels=5, output_shape=1, convolution_bond=3)
    record("lanczos", "conv_lanczos_reg", layer, X, y, SquareBregFunction(), batch_size=64, num_swipes=2, lr=1.0, max_iter=5, tol=1e-12)
    for solver in ("cg", "minres"):
        torch.manual_seed(7)
        X, y = data(3, 140, 5, 4, K=4)
        layer = TensorConvolutionTrainLayer(num_carriages=3, bond_dim=4, num_patches=5, patch_pixels=4, output_shape=3, convolution_bond=2)
        record("scipy", f"conv_scipy_{solver}", layer, X, y, XEAutogradBregman(w=1.0), solver=solver, batch_size=70, num_swipes=2, lr=1.0,
               max_iter=25, tol=1e-5)
    torch.manual_seed(8)
    X, y = data(4, 100, 5, 4, C=1)
    layer = TensorConvolutionTrainLayer(num_carriages=2, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=1, convolution_bond=2)
    record("scipy", "conv_scipy_cg_2col", layer, X, y, SquareBregFunction(), solver="cg", batch_size=-1, num_swipes=1, lr=1.0, max_iter=25, tol=1e-5)


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def record_dense(name, layer, X, y, loss_fn, **kw):
    """accumulating_swipe on the conv layer (image_convolution_MNIST.py:120 call shape)."""
    tn = layer.tensor_network
    cores0 = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    names = [n.name for n in tn.train_nodes]
    pred0 = tn.forward(X, to_tensor=True).detach().numpy().copy()
    tn.reset_stacks()
    ups, losses = [], []

    def block_callback(NS, node):
        ups.append({"NS": NS, "k": tn.train_nodes.index(node), "after": [n.tensor.detach().numpy().copy() for n in tn.train_nodes]})

    ok = tn.accumulating_swipe(X, y, loss_fn, block_callback=block_callback, loss_callback=lambda NS, node, l: losses.append(float(l)), **kw)
    tn.reset_stacks()
    pred1 = tn.forward(X, to_tensor=True).detach().numpy().copy()
    flat = {"x": X.numpy(), "y": y.numpy(), "n_cores": np.array(len(cores0)), "n_updates": np.array(len(ups)), "ok": np.array(bool(ok)),
            "losses": np.array(losses), "names": np.array(names), "pred0": pred0, "pred1": pred1}
    for i, c in enumerate(cores0):
        flat[f"cores0_{i}"] = c
    for ui, u in enumerate(ups):
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"]])
        for i, c in enumerate(u["after"]):
            flat[f"u{ui}_after_{i}"] = c
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print(name, names, len(ups), "updates", "ok", ok, "losses", losses[:3], "->", losses[-1])


def main_dense():
    torch.manual_seed(15)
    X, y = data(5, 130, 5, 4, K=3)
    layer = TensorConvolutionTrainLayer(num_carriages=3, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2)
    record_dense("conv_dense_xe", layer, X, y, XEAutogradBregman(w=1.0), batch_size=50, num_swipes=1, lr=1.0, method="ridge_exact", eps=1.0,
                 eps_decay=0.5)
    torch.manual_seed(16)
    X, y = data(6, 140, 6, 5, C=1)
    layer = TensorConvolutionTrainLayer(num_carriages=3, bond_dim=3, num_patches=6, patch_pixels=5, output_shape=1, convolution_bond=2)
    record_dense("conv_dense_reg", layer, X, y, SquareBregFunction(), batch_size=-1, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5,
                 eps_decay=0.7)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dense":
    main_dense()
