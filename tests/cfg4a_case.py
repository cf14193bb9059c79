"""Shared body of the BASELINE config-4a test (TNML classifier, sin-cos map, 9 logits, cross-entropy, rank 38 with local systems up to
2888 parameters, scipy_swipe(cg)) on a 16-site chain: tests/golden/make_golden_cfg4a.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

N, SITES, R, C = 512, 16, 38, 9


def data():
    rng = np.random.default_rng(2030)
    X = rng.uniform(0, 1, size=(N, SITES))
    y = np.eye(C + 1)[np.argmax(X @ rng.normal(size=(SITES, C + 1)), axis=1)]
    return X, y


def run(device, fused_map=True):
    """(relative errors of the 32 per-node losses, prediction error on the first 128 rows); the SciPy solver object is passed, so the
    Krylov recurrences run in float32 on the host as in the reference (network.py:918-926)."""
    from scipy.sparse.linalg import cg
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg4a_chain16.npz"))
    X, y = data()
    assert np.array_equal(X[:2], z["x_head"])
    layer = tnb.TensorTrainLayer(SITES, R, 2, output_shape=C, constrict_bond=True, seed=42)
    layer.to(device)
    tn = layer.tensor_network
    Xt = torch.tensor(X, device=device)
    if fused_map:
        x = tnb.MappedInput(Xt, "sin-cos")
    else:
        x = [torch.stack([torch.cos(0.5 * np.pi * Xt[:, j]), torch.sin(0.5 * np.pi * Xt[:, j])], 1) for j in range(SITES)]
    losses = []
    ok = tn.scipy_swipe(x, torch.tensor(y, device=device), tnb.XEAutogradBregman(w=1.0), cg, batch_size=512, num_swipes=2, lr=0.05, max_iter=5,
                        tol=1e-3, loss_callback=lambda l: losses.append(float(l)))
    assert ok == bool(z["ok"]) and len(losses) == len(z["losses"])
    loss_err = np.abs(np.array(losses) - z["losses"]) / np.abs(z["losses"])
    xs = x[:128] if fused_map else [t[:128] for t in x]
    pred = tn.forward(xs, to_tensor=True).cpu().numpy()
    return loss_err, gu.relerr(pred, z["pred128"])
