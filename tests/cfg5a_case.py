"""Shared body of the BASELINE config-5a Gram test (TT poly-mode, 5 cores, rank 38, 28 features + bias: P = 41 876 at the middle core)
against the fingerprint of the reference's own get_A_b: tests/golden/make_golden_cfg5a.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb
from tensornetworksfork_b200 import ops

N, F, R, CORES, SITE = 512, 28, 38, 5, 2


def data():
    rng = np.random.default_rng(2029)
    X = rng.uniform(-1, 1, size=(N, F))
    W = rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1))
    return np.concatenate([X, np.ones((N, 1))], 1), y


def setup(device, gram_mode="fp64"):
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg5a_gram.npz"))
    X, y = data()
    assert np.array_equal(X[:2], z["x_head"])
    layer = tnb.TensorTrainLayer(CORES, R, F + 1, output_shape=1, constrict_bond=False, seed=42)
    tn = layer.tensor_network
    sums = np.array([float(n.tensor.abs().sum()) for n in tn.train_nodes])
    assert np.array_equal(sums, z["core_abs_sums"]), "the constructor did not reproduce the reference's initial cores"
    layer.to(device)
    tn.gram_mode = gram_mode
    Xt, yt = torch.tensor(X, device=device), torch.tensor(y, device=device)
    tn.set_input(Xt)
    tn._check_external()
    return z, tn, Xt, yt


def matrix_free(device):
    """Errors of (prediction, per-row loss, b, A v) with A v = J^T diag(w) (J v) from the engine's factors -- no P x P matrix."""
    z, tn, Xt, yt = setup(device)
    prob = tn._site_problem(SITE, yt, tnb.SquareBregFunction())
    rf, gf = prob["rhs"], prob["gram"]
    b = ops.rhs(rf[0], rf[1], rf[2], prob["rw"], prob["rrows"])
    v = torch.tensor(np.random.default_rng(7).normal(size=b.numel()), device=device)
    Av = ops.matvec(gf[0], gf[1], gf[2], prob["gw"], prob["grows"], v)
    return (gu.relerr(prob["yhat"].cpu().numpy(), z["pred"]), gu.relerr(prob["loss"].cpu().numpy().reshape(z["loss"].shape), z["loss"]),
            gu.relerr(b.cpu().numpy(), z["b"]), gu.relerr(Av.cpu().numpy(), z["Av"]))


def dense(device, gram_mode):
    """Errors of (b, diag A, A v, Frobenius norm, asymmetry) of the dense 41 876 x 41 876 matrix the engine expands from its unique
    entries (14 GB: GPU only)."""
    z, tn, Xt, yt = setup(device, gram_mode)
    A, b = tn.get_A_b(tn.train_nodes[SITE], y=yt, loss_fn=tnb.SquareBregFunction())
    P = b.numel()
    A = A.reshape(P, P)
    v = torch.tensor(np.random.default_rng(7).normal(size=P), device=device)
    out = (gu.relerr(b.reshape(P).cpu().numpy(), z["b"]), gu.relerr(A.diagonal().cpu().numpy(), z["diagA"]),
           gu.relerr((A @ v).cpu().numpy(), z["Av"]), abs(float(torch.linalg.matrix_norm(A)) - float(z["fro"])) / float(z["fro"]),
           float((A - A.t()).abs().max()))
    return out
