"""Shared body of the BASELINE config-5b chain test (28-site TNML, polynomial basis of degree 5, rank 38, QR re-gauge; largest local
system P = 8664): tests/golden/make_golden_cfg5b.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

N, F, R, DEG = 2048, 28, 38, 5


def data():
    rng = np.random.default_rng(2027)
    X = rng.uniform(-1, 1, size=(N, F))
    W = rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1))
    return X, y


def run(device, max_updates=None, gram_mode="fp64"):
    """Relative loss errors of the first ``max_updates`` (all 55 when None) updates and, for a full sweep, the final prediction error."""
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg5b_chain28.npz"))
    X, y = data()
    assert np.array_equal(X[:4], z["x_head"]) and np.array_equal(y[:4], z["y_head"])
    layer = tnb.TensorTrainLayer(F, R, DEG + 1, output_shape=1, constrict_bond=True, seed=42)
    layer.to(device)
    tn = layer.tensor_network
    tn.gram_mode = gram_mode
    tn.small_site_fp64 = 0          # exercise the requested Gram mode even where the site is small
    x = tnb.MappedInput(torch.tensor(X, device=device), "polynomial", degree=DEG)
    tn.orthonormalize_left()
    trace = []

    def stop():
        return max_updates is not None and len(trace) >= max_updates

    ok = tn.accumulating_swipe(x, torch.tensor(y, device=device), tnb.SquareBregFunction(), batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5,
                               orthonormalize=True, method="ridge_cholesky", num_swipes=1, skip_second=False, direction="l2r",
                               convergence_criterion=stop, loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    assert ok
    ref = z["trace"][:len(trace)]
    assert [(a, b) for a, b, _ in trace] == [(int(a), int(b)) for a, b, _ in ref]
    loss_err = np.array([abs(t[2] - r[2]) / max(abs(r[2]), 1e-300) for t, r in zip(trace, ref)])
    pred_err = None
    if max_updates is None:
        pred = tn.forward(x[:256], to_tensor=True).cpu().numpy()
        pred_err = gu.relerr(pred.reshape(z["pred256"].shape), z["pred256"])
    return loss_err, pred_err
