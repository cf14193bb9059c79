"""Shared body of the full-size BASELINE config-2 parity test (CPD rank 100): tests/golden/make_golden_cfg2.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

N, F, RANK, FACTORS, NUM_SWIPES = 20640, 8, 100, 5, 2


def data():
    rng = np.random.default_rng(2025)
    X = rng.uniform(-1, 1, size=(N, F))
    W1, W2 = rng.normal(size=(F, 1)) / np.sqrt(F), rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W1) + 0.5 * (X @ W2) ** 2 + 0.3 * X[:, :1] * X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1))
    return np.concatenate([X, np.ones((N, 1))], 1), y


def load():
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg2_full.npz"))
    X, y = data()
    assert np.array_equal(X[:4], z["x_head"]) and np.array_equal(y[:4], z["y_head"])
    return z, X, y


def run(device, gram_mode="fp64"):
    """Per-update relative loss errors and the final prediction error on the first 256 rows."""
    z, X, y = load()
    layer = tnb.CPDLayer(FACTORS, RANK, F + 1, output_shape=(1,), seed=42)
    layer.to(device)
    tn = layer.tensor_network
    tn.gram_mode = gram_mode
    tn.small_site_fp64 = 0          # exercise the requested Gram mode even where the site is small
    trace = []
    ok = tn.accumulating_swipe(torch.tensor(X, device=device), torch.tensor(y, device=device), tnb.SquareBregFunction(), batch_size=512, lr=1.0,
                               eps=1.0, eps_decay=0.5, orthonormalize=False, method="ridge_cholesky", num_swipes=NUM_SWIPES, skip_second=False,
                               direction="l2r", loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    ref = z["trace"]
    assert ok == bool(z["ok"])
    assert [(a, b) for a, b, _ in trace] == [(int(a), int(b)) for a, b, _ in ref]
    loss_err = np.array([abs(t[2] - r[2]) / max(abs(r[2]), 1e-300) for t, r in zip(trace, ref)])
    pred = tn.forward(torch.tensor(X[:256], device=device), to_tensor=True).cpu().numpy()
    return loss_err, gu.relerr(pred.reshape(z["pred256"].shape), z["pred256"])
