"""Shared body of the BASELINE config-4b test (conv-TT at full model size: 50 patches x 17 pixels, r = 38, CB = 4, 9 logits; patch cores
of 17 100 / 72 200 / 1 900 parameters) under scipy_swipe(minres): tests/golden/make_golden_cfg4b.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

N, Q, T, R, CB, C = 256, 50, 17, 38, 4, 9


def data():
    rng = np.random.default_rng(2028)
    X = rng.uniform(0, 1, size=(N, Q, T))
    X[:, -1, :] = 0.0
    X[:, :, -1] = 0.0
    X[:, -1, -1] = 1.0
    y = np.eye(C + 1)[rng.integers(0, C + 1, N)]
    return X, y


def run(device):
    """(relative errors of the six per-node losses, prediction error on the first 64 rows); the SciPy solver object is passed, so the
    Krylov recurrences run in float32 on the host exactly as in the reference (network.py:918-926)."""
    from scipy.sparse.linalg import minres
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg4b_shape.npz"))
    X, y = data()
    assert np.array_equal(X[:2], z["x_head"])
    torch.manual_seed(42)
    layer = tnb.TensorConvolutionTrainLayer(num_carriages=3, bond_dim=R, num_patches=Q, patch_pixels=T, output_shape=C, convolution_bond=CB)
    tn = layer.tensor_network
    assert [n.name for n in tn.train_nodes] == [str(s) for s in z["names"]]
    sums = np.array([float(n.tensor.double().abs().sum()) for n in tn.train_nodes])
    assert np.array_equal(sums, z["core_abs_sums0"]), "the constructor did not reproduce the reference's initial cores"
    layer.to(device)
    losses = []
    ok = tn.scipy_swipe(torch.tensor(X, device=device), torch.tensor(y, device=device), tnb.XEAutogradBregman(w=1.0), minres, batch_size=256,
                        num_swipes=1, lr=1.0, max_iter=3, tol=1e-3, loss_callback=lambda l: losses.append(float(l)))
    assert ok == bool(z["ok"]) and len(losses) == len(z["losses"])
    loss_err = np.abs(np.array(losses) - z["losses"]) / np.abs(z["losses"])
    pred = tn.forward(torch.tensor(X[:64], device=device), to_tensor=True).cpu().numpy()
    return loss_err, gu.relerr(pred, z["pred64"])
