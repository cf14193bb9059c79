"""Shared body of the conv-TT (patch/pixel) parity tests: CPU stand-in kernels and real kernels on the GPU."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

CASES = {
    "conv_lanczos_xe": dict(kind="lanczos", ctor=dict(num_carriages=3, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2),
                            loss=lambda: tnb.XEAutogradBregman(w=1.0), oloss="xe",
                            kw=dict(batch_size=50, num_swipes=2, lr=1.0, max_iter=6, tol=1e-12)),
    "conv_lanczos_reg": dict(kind="lanczos", ctor=dict(num_carriages=4, bond_dim=3, num_patches=6, patch_pixels=5, output_shape=1, convolution_bond=3),
                             loss=lambda: tnb.SquareBregFunction(), oloss="square",
                             kw=dict(batch_size=64, num_swipes=2, lr=1.0, max_iter=5, tol=1e-12)),
    "conv_scipy_cg": dict(kind="scipy", solver="cg", ctor=dict(num_carriages=3, bond_dim=4, num_patches=5, patch_pixels=4, output_shape=3, convolution_bond=2),
                          loss=lambda: tnb.XEAutogradBregman(w=1.0), oloss="xe",
                          kw=dict(batch_size=70, num_swipes=2, lr=1.0, max_iter=25, tol=1e-5)),
    "conv_scipy_minres": dict(kind="scipy", solver="minres", ctor=dict(num_carriages=3, bond_dim=4, num_patches=5, patch_pixels=4, output_shape=3, convolution_bond=2),
                              loss=lambda: tnb.XEAutogradBregman(w=1.0), oloss="xe",
                              kw=dict(batch_size=70, num_swipes=2, lr=1.0, max_iter=25, tol=1e-5)),
    "conv_scipy_cg_2col": dict(kind="scipy", solver="cg", ctor=dict(num_carriages=2, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=1, convolution_bond=2),
                               loss=lambda: tnb.SquareBregFunction(), oloss="square",
                               kw=dict(batch_size=-1, num_swipes=1, lr=1.0, max_iter=25, tol=1e-5)),
    "conv_dense_xe": dict(kind="dense", ctor=dict(num_carriages=3, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2),
                          loss=lambda: tnb.XEAutogradBregman(w=1.0), oloss="xe",
                          kw=dict(batch_size=50, num_swipes=1, lr=1.0, method="ridge_exact", eps=1.0, eps_decay=0.5)),
    "conv_dense_reg": dict(kind="dense", ctor=dict(num_carriages=3, bond_dim=3, num_patches=6, patch_pixels=5, output_shape=1, convolution_bond=2),
                           loss=lambda: tnb.SquareBregFunction(), oloss="square",
                           kw=dict(batch_size=-1, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5, eps_decay=0.7)),
}


def load(name):
    fx = gu.load_krylov(name)
    z = np.load(os.path.join(gu.GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    fx["names"] = [str(s) for s in z["names"]]
    fx["pred0"], fx["pred1"] = z["pred0"], z["pred1"]
    return fx


def build(name, device, chunk_rows=None):
    case = CASES[name]
    fx = load(name)
    layer = tnb.TensorConvolutionTrainLayer(**case["ctor"])
    tn = layer.tensor_network
    assert [n.name for n in tn.train_nodes] == fx["names"]
    for n, c in zip(tn.train_nodes, fx["cores0"]):
        assert tuple(n.tensor.shape) == c.shape, (n.name, n.tensor.shape, c.shape)
        n.tensor = torch.tensor(c, device=device)
    if chunk_rows is not None:
        tn.chunk_rows = chunk_rows
    return case, fx, layer


def run_case(name, device, scipy_object=True, chunk_rows=None, loss_prefix=None):
    """Returns (forward error, max relative core error over all updates, max loss error, final prediction error)."""
    case, fx, layer = build(name, device, chunk_rows)
    tn = layer.tensor_network
    X = torch.tensor(fx["x"], device=device)
    y = torch.tensor(fx["y"], device=device)
    fwd_err = gu.relerr(layer(X).cpu().numpy(), fx["pred0"])
    ups, losses = [], []

    def block_callback(NS, node):
        ups.append((NS, tn.train_nodes.index(node), [n.tensor.cpu().numpy().copy() for n in tn.train_nodes]))

    if case["kind"] == "dense":
        ok = tn.accumulating_swipe(X, y, case["loss"](), block_callback=block_callback,
                                   loss_callback=lambda NS, node, l: losses.append(l), **case["kw"])
    elif case["kind"] == "lanczos":
        x0s = [u["x0"] for u in fx["updates"]]
        cnt = [0]

        def x0_fn(node, b):
            v = torch.tensor(x0s[cnt[0]], device=device)
            cnt[0] += 1
            return v.reshape(-1)

        ok = tn.lanczos_swipe(X, y, case["loss"](), block_callback=block_callback, loss_callback=losses.append, x0_fn=x0_fn, **case["kw"])
    else:
        if scipy_object:
            from scipy.sparse.linalg import cg, minres
            solver = {"cg": cg, "minres": minres}[case["solver"]]
        else:
            solver = case["solver"]
        ok = tn.scipy_swipe(X, y, case["loss"](), solver, block_callback=block_callback, loss_callback=losses.append, **case["kw"])
    assert ok
    assert [(a, b) for a, b, _ in ups] == [(u["NS"], u["k"]) for u in fx["updates"]]
    core_err = 0.0
    for (_, _, cores), u in zip(ups, fx["updates"]):
        for c, ref in zip(cores, u["after"]):
            core_err = max(core_err, gu.relerr(c, ref))
    le = np.abs(np.array(losses) - fx["losses"]) / np.maximum(1.0, np.abs(fx["losses"]))
    loss_err = float(np.max(le[:loss_prefix] if loss_prefix else le))
    pred_err = gu.relerr(layer(X).cpu().numpy(), fx["pred1"])
    return fwd_err, core_err, loss_err, pred_err
