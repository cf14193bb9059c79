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
    # type-I image model (AAMNST.py:157-168): sum of conv-TTs with 1..3 columns, members after the first without bias patch / pixel
    "conv_type1": dict(kind="dense", builder=lambda: _conv_type1(), loss=lambda: tnb.XEAutogradBregman(w=1.0), oloss="xe",
                       kw=dict(batch_size=50, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)),
    "conv_onecol": dict(kind="dense", ctor=dict(num_carriages=1, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2),
                        loss=lambda: tnb.SquareBregFunction(), oloss="square",
                        kw=dict(batch_size=40, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5, eps_decay=0.5)),
    "conv_nocb": dict(kind="dense", ctor=dict(num_carriages=3, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=-1),
                      loss=lambda: tnb.XEAutogradBregman(w=1.0), oloss="xe",
                      kw=dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_exact", eps=0.8, eps_decay=0.5)),
}


def _conv_type1():
    nets = [tnb.TensorConvolutionTrainLayer(num_carriages=i, bond_dim=3, num_patches=5 if i == 1 else 4, patch_pixels=4 if i == 1 else 3,
                                            output_shape=2, convolution_bond=2).tensor_network for i in range(1, 4)]
    return tnb.TensorNetworkLayer(tnb.SumOfNetworks(nets, train_operators=True))


def load(name):
    fx = gu.load_krylov(name)
    z = np.load(os.path.join(gu.GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    fx["names"] = [str(s) for s in z["names"]]
    fx["pred0"], fx["pred1"] = z["pred0"], z["pred1"]
    return fx


def build(name, device, chunk_rows=None):
    case = CASES[name]
    fx = load(name)
    layer = case["builder"]() if "builder" in case else tnb.TensorConvolutionTrainLayer(**case["ctor"])
    tn = layer.tensor_network
    assert [n.name for n in tn.train_nodes] == fx["names"]
    for n, c in zip(tn.train_nodes, fx["cores0"]):
        assert tuple(n.tensor.shape) == c.shape, (n.name, n.tensor.shape, c.shape)
        n.tensor = torch.tensor(c, device=device)
    if chunk_rows is not None:
        for net in getattr(tn, "networks", [tn]):
            net.chunk_rows = chunk_rows
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


# ---------------------------------------------------------------------------------------------------------------------------
# growing flow (grow_cart): tests/golden/make_golden_conv_grow.py
GROW_CTOR = dict(num_carriages=2, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2)
GROW_PHASES = [
    dict(grow=None, kw=dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_exact", eps=[1.0, 0.5])),
    dict(grow=(3, 2), kw=dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_exact", eps=[0.8, 0.4], direction="r2l")),
    dict(grow=(None, None), kw=dict(batch_size=48, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.6, eps_decay=0.5)),
]


def load_grow():
    return np.load(os.path.join(gu.GOLDEN_DIR, "conv_grow.npz"), allow_pickle=False)


def run_grow(device):
    """The three phases of the recording on one layer; per phase (forward error after the growth, max relative core error over
    all updates, max loss error, final prediction error).  After each growth the cores the growth touched (broadcast old last
    cores, constant new patch core) must agree with the reference's; the freshly drawn pixel core is taken from the recording
    (its draw depends on the RNG state; test_grow_cart_reproduces_reference_cores_bit_for_bit covers the draw itself)."""
    z = load_grow()
    X = torch.tensor(z["x"], device=device)
    y = torch.tensor(z["y"], device=device)
    layer = tnb.TensorConvolutionTrainLayer(**GROW_CTOR)
    out = []
    for pi, ph in enumerate(GROW_PHASES):
        pre = f"p{pi}_"
        if ph["grow"] is not None:
            layer.grow_cart(*ph["grow"])
        tn = layer.tensor_network
        names = [str(s) for s in z[pre + "names"]]
        assert [n.name for n in tn.train_nodes] == names
        fresh = names[-2] if ph["grow"] is not None else None          # the new pixel core (train-node order ..., C_new, A_new)
        for i, n in enumerate(tn.train_nodes):
            ref = z[pre + f"cores0_{i}"]
            assert tuple(n.tensor.shape) == ref.shape, (n.name, tuple(n.tensor.shape), ref.shape)
            if pi == 0 or n.name == fresh:
                n.tensor = torch.tensor(ref, device=device)
            elif ph["grow"] is not None and n.name in names[-4:]:
                # cores touched by the growth: broadcast old last cores and the constant new patch core
                assert gu.relerr(n.tensor.cpu().numpy(), ref) < 1e-7, n.name
        if device != "cpu":
            layer.cuda()
        fwd_err = gu.relerr(layer(X).cpu().numpy(), z[pre + "pred0"])
        ups, losses = [], []

        def block_callback(NS, node, tn=tn, ups=ups):
            ups.append((NS, tn.train_nodes.index(node), [n.tensor.cpu().numpy().copy() for n in tn.train_nodes]))

        ok = tn.accumulating_swipe(X, y, tnb.XEAutogradBregman(w=1.0), block_callback=block_callback,
                                   loss_callback=lambda NS, node, l, losses=losses: losses.append(float(l)), **ph["kw"])
        assert ok
        nu = int(z[pre + "n_updates"])
        assert [(a, b) for a, b, _ in ups] == [tuple(int(v) for v in z[pre + f"u{ui}_scal"]) for ui in range(nu)]
        core_err = 0.0
        for ui, (_, _, cores) in enumerate(ups):
            for i, c in enumerate(cores):
                core_err = max(core_err, gu.relerr(c, z[pre + f"u{ui}_after_{i}"]))
        ref_losses = z[pre + "losses"]
        loss_err = float(np.max(np.abs(np.array(losses) - ref_losses) / np.maximum(1.0, np.abs(ref_losses))))
        pred_err = gu.relerr(layer(X).cpu().numpy(), z[pre + "pred1"])
        out.append((fwd_err, core_err, loss_err, pred_err))
    return out
