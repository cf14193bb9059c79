"""Shared body of the cum-sum train parity test (CPU stand-in kernels and real kernels)."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb


def run(device, teacher_forced=True):
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cumsum_reg.npz"))
    n = int(z["n_cores"])
    X = torch.tensor(z["x"], device=device)
    y = torch.tensor(z["y"], device=device)
    layer = tnb.CumSumLayer(n, 3, 4, output_shape=1, constrict_bond=False)
    tn = layer.tensor_network
    for i, nd in enumerate(tn.train_nodes):
        nd.tensor = torch.tensor(z[f"cores0_{i}"], device=device)
    pred0 = tn.forward(X, to_tensor=True).cpu().numpy()
    assert gu.relerr(pred0.reshape(z["pred0"].shape), z["pred0"]) < 1e-12        # closed form == the reference's operator nodes
    trace = []
    ok = tn.accumulating_swipe(X, y, tnb.SquareBregFunction(), batch_size=64, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5,
                               eps_decay=0.5, loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), l)))
    assert ok
    nu = int(z["n_updates"])
    assert len(trace) == nu
    for ui, (NS, k, l) in enumerate(trace):
        rNS, rk, rl = z[f"u{ui}_scal"]
        assert (NS, k) == (int(rNS), int(rk))
        assert abs(l - rl) <= 1e-7 * max(1.0, abs(rl)), (ui, l, rl)
    pred = tn.forward_batch(X, 64).cpu().numpy()
    assert gu.relerr(pred.reshape(z["pred"].shape), z["pred"]) < 1e-7
    for i, nd in enumerate(tn.train_nodes):
        assert gu.relerr(nd.tensor.cpu().numpy(), z[f"u{nu - 1}_after_{i}"]) < 1e-6
    # teacher-forced Gram / rhs of the middle core against the recorded dense A, b
    for ui in range(nu if teacher_forced else 0):
        k = int(z[f"u{ui}_scal"][1])
        for i, nd in enumerate(tn.train_nodes):
            nd.tensor = torch.tensor(z[f"u{ui}_before_{i}"], device=device)
        tn.set_input(X)
        tn._check_external()
        A, b = cumsum_A_b(tn, k, y)
        P = b.numel()
        assert gu.relerr(A.cpu().numpy().reshape(P, P), z[f"u{ui}_A"].reshape(P, P)) < 1e-12
        assert gu.relerr(b.cpu().numpy().ravel(), z[f"u{ui}_b"].ravel()) < 1e-12


def cumsum_A_b(tn, k, y):
    """Dense A, b of core k as the engine builds them (without solving)."""
    from tensornetworksfork_b200 import ops
    from tensornetworksfork_b200.ops import Factor
    from tensornetworksfork_b200.tensor.bregman import hessian_terms
    _, facs, S, dev = tn._data
    G = tn._canon(k)
    rl, _, f, rr = G.shape
    L, R = tn._get_left(k - 1), tn._get_right(k + 1)
    yhat = tn._predict_at(k, L, R, facs[k], S)
    loss, g, U, lam = hessian_terms(tnb.SquareBregFunction(), yhat, y)
    w = (lam.reshape(S, -1) * U.reshape(S, -1) ** 2).sum(1).contiguous()
    one = ops.ones_factor(G)
    f1 = one if L is None else Factor(L.reshape(S, f * rl), m=f * rl)
    f3 = one if R is None else Factor(R.reshape(S, f * rr), m=f * rr)
    t1, t2, t3 = tn._tables(rl, f, rr, L is not None, R is not None, dev)
    A = ops.gram_generic(f1, facs[k], f3, t1, t2, t3, w, S)
    b = ops.gram_generic(f1, facs[k], f3, t1, t2, t3, g.reshape(S).contiguous(), S, rhs_only=True)
    return A, b
