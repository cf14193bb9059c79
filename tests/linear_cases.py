"""Shared body of the TensorTrainLinearLayer parity tests (CPU stand-in kernels / real kernels)."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

CASES = {
    "linear_tt_reg": dict(ctor=(3, 3, 6, 4), out=1, seed=5, loss=lambda: tnb.SquareBregFunction(), kind="dense",
                          kw=dict(batch_size=100, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)),
    "linear_tt_xe": dict(ctor=(3, 3, 5, 3), out=2, seed=6, loss=lambda: tnb.XEAutogradBregman(w=1.0), kind="dense",
                         kw=dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.5)),
    "linear_tt_lanczos": dict(ctor=(3, 3, 5, 3), out=1, seed=7, loss=lambda: tnb.SquareBregFunction(), kind="lanczos",
                              kw=dict(batch_size=80, num_swipes=2, lr=1.0, max_iter=5, tol=1e-12)),
}


def run_case(name, device):
    """(initial-core error, forward error, max core error over all updates, max loss error, final prediction error)."""
    case = CASES[name]
    z = np.load(os.path.join(gu.GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    fx = gu.load_krylov(name)
    layer = tnb.TensorTrainLinearLayer(*case["ctor"], output_shape=case["out"], constrict_bond=False, seed=case["seed"])
    tn = layer.tensor_network
    assert [n.name for n in tn.train_nodes] == [str(s) for s in z["names"]]
    init_err = max(gu.relerr(n.tensor.numpy(), c) for n, c in zip(tn.train_nodes, fx["cores0"]))     # same seed, same draws
    layer.to(device)
    X, y = torch.tensor(fx["x"], device=device), torch.tensor(fx["y"], device=device)
    fwd_err = gu.relerr(layer(X).cpu().numpy().reshape(z["pred0"].shape), z["pred0"])
    ups, losses = [], []

    def block_callback(NS, node):
        ups.append((NS, tn.train_nodes.index(node), [n.tensor.cpu().numpy().copy() for n in tn.train_nodes]))

    if case["kind"] == "dense":
        ok = tn.accumulating_swipe(X, y, case["loss"](), block_callback=block_callback,
                                   loss_callback=lambda NS, node, l: losses.append(l), **case["kw"])
    else:
        x0s = [u["x0"] for u in fx["updates"]]
        cnt = [0]

        def x0_fn(node, b):
            v = torch.tensor(x0s[cnt[0]], device=device)
            cnt[0] += 1
            return v.reshape(-1)

        ok = tn.lanczos_swipe(X, y, case["loss"](), block_callback=block_callback, loss_callback=losses.append, x0_fn=x0_fn, **case["kw"])
    assert ok
    assert [(a, b) for a, b, _ in ups] == [(u["NS"], u["k"]) for u in fx["updates"]]
    per_update = [max(gu.relerr(c, ref) for c, ref in zip(cores, u["after"])) for (_, _, cores), u in zip(ups, fx["updates"])]
    # the free-running Lanczos sweep (no ridge, gauge-singular systems) amplifies rounding differences update by update: the
    # first pass over the six nodes is compared tightly, the whole run loosely
    core_err = max(per_update[:6]) if case["kind"] == "lanczos" else max(per_update)
    assert max(per_update) < 1e-4, per_update
    loss_err = float(np.max(np.abs(np.array(losses) - fx["losses"]) / np.maximum(1.0, np.abs(fx["losses"]))))
    pred_err = gu.relerr(layer(X).cpu().numpy().reshape(z["pred1"].shape), z["pred1"])
    return init_err, fwd_err, core_err, loss_err, pred_err
