"""Shared body of the BASELINE config-3 chain test (90-site TNML, sin-cos map, rank 24, QR re-gauge): tests/golden/make_golden_cfg3.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

N, F, R = 4096, 90, 24


def data():
    rng = np.random.default_rng(2026)
    X = rng.uniform(-1, 1, size=(N, F))
    W = rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1))
    return X, y


def run(device, max_updates=None, fused_map=True, gram_mode="fp64"):
    """Relative loss errors of the first ``max_updates`` (all 179 when None) updates and, for a full sweep, the final prediction error.
    ``fused_map``: hand the engine the raw matrix + feature map (MappedInput, what TNMLRegressor does) instead of 90 mapped tensors."""
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg3_chain90.npz"))
    X, y = data()
    assert np.array_equal(X[:4], z["x_head"]) and np.array_equal(y[:4], z["y_head"])
    layer = tnb.TensorTrainLayer(F, R, 2, output_shape=1, constrict_bond=True, seed=42)
    layer.to(device)
    tn = layer.tensor_network
    tn.gram_mode = gram_mode
    Xt = torch.tensor(X, device=device)
    if fused_map:
        x = tnb.MappedInput(Xt, "sin-cos")
    else:
        x = [torch.stack([torch.cos(0.5 * np.pi * Xt[:, j]), torch.sin(0.5 * np.pi * Xt[:, j])], 1) for j in range(F)]
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
        xs = x[:256] if fused_map else [t[:256] for t in x]
        pred = tn.forward(xs, to_tensor=True).cpu().numpy()
        pred_err = gu.relerr(pred.reshape(z["pred256"].shape), z["pred256"])
    return loss_err, pred_err
