"""Shared body of the matrix-free sweep parity tests (CPU stand-in kernels and real kernels on the GPU)."""
import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

CASES = {
    "krylov_lanczos_reg": dict(kind="lanczos", n=4, r=3, f=4, C=1, loss=lambda: tnb.SquareBregFunction(),
                               kw=dict(batch_size=50, num_swipes=2, lr=1.0, max_iter=6, tol=1e-12)),
    "krylov_lanczos_xe": dict(kind="lanczos", n=3, r=3, f=4, C=2, loss=lambda: tnb.XEAutogradBregman(w=1.0),
                              kw=dict(batch_size=80, num_swipes=1, lr=1.0, max_iter=5, tol=1e-12)),
    "krylov_scipy_cg": dict(kind="scipy", solver="cg", n=4, r=3, f=4, C=1, loss=lambda: tnb.SquareBregFunction(),
                            kw=dict(batch_size=60, num_swipes=2, lr=1.0, max_iter=25, tol=1e-5)),
    "krylov_scipy_minres": dict(kind="scipy", solver="minres", n=4, r=3, f=4, C=1, loss=lambda: tnb.SquareBregFunction(),
                                kw=dict(batch_size=60, num_swipes=2, lr=1.0, max_iter=25, tol=1e-5)),
    # cum-sum train (CumSumLayer): J v = prediction with v in place of the core, J^T u = table-driven right-hand-side pass
    "krylov_cumsum_lanczos": dict(kind="lanczos", cumsum=True, n=3, r=3, f=4, C=1, loss=lambda: tnb.SquareBregFunction(),
                                  kw=dict(batch_size=60, num_swipes=2, lr=1.0, max_iter=6, tol=1e-12)),
    "krylov_cumsum_cg": dict(kind="scipy", solver="cg", cumsum=True, n=4, r=2, f=4, C=1, loss=lambda: tnb.SquareBregFunction(),
                             kw=dict(batch_size=-1, num_swipes=2, lr=1.0, max_iter=8, tol=1e-5)),
}


def run_case(name, device, scipy_object=True):
    """Returns (max relative core error over all updates, max loss error) against the reference recording."""
    case = CASES[name]
    fx = gu.load_krylov(name)
    if case.get("cumsum"):
        layer = tnb.CumSumLayer(case["n"], case["r"], case["f"], output_shape=case["C"], constrict_bond=False, perturb=False)
    else:
        layer = tnb.TensorTrainLayer(case["n"], case["r"], case["f"], output_shape=case["C"], constrict_bond=False, seed=0)
    tn = layer.tensor_network
    assert [tuple(n.tensor.shape) for n in tn.train_nodes] == [c.shape for c in fx["cores0"]]
    for n, c in zip(tn.train_nodes, fx["cores0"]):
        n.tensor = torch.tensor(c, device=device)
    X = torch.tensor(fx["x"], device=device)
    y = torch.tensor(fx["y"], device=device)
    ups = []
    losses = []
    it = iter(fx["updates"])

    def block_callback(NS, node):
        ups.append((NS, tn.train_nodes.index(node), [n.tensor.cpu().numpy().copy() for n in tn.train_nodes]))

    if case["kind"] == "lanczos":
        x0s = [u["x0"] for u in fx["updates"]]
        cnt = [0]

        def x0_fn(node, b):
            v = torch.tensor(x0s[cnt[0]], device=device)
            cnt[0] += 1
            # the recording is in the node's own layout; the engine works in canonical (a, c, p, b) order = same here
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
    loss_err = float(np.max(np.abs(np.array(losses) - fx["losses"]) / np.maximum(1.0, np.abs(fx["losses"]))))
    return core_err, loss_err
