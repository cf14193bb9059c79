"""Shared body of the full-size BASELINE config-1 parity test: tests/golden/make_golden_cfg1.py."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb

N, F, R, NUM_SWIPES = 4177, 8, 6, 4
EPSS = np.geomspace(0.07542717629430484, 0.00000000000722857583, 2 * NUM_SWIPES).tolist()


def data():
    rng = np.random.default_rng(2024)
    X = rng.uniform(-1, 1, size=(N, F))
    W1, W2 = rng.normal(size=(F, 1)) / np.sqrt(F), rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W1) + 0.5 * (X @ W2) ** 2 + 0.3 * X[:, :1] * X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1))
    return np.concatenate([X, np.ones((N, 1))], 1), y


def run(device, gram_mode="fp64"):
    """Per-update relative loss errors (array), final prediction error on the first 256 rows, max relative core error."""
    z = np.load(os.path.join(gu.GOLDEN_DIR, "cfg1_full.npz"))
    X, y = data()
    assert np.array_equal(X[:4], z["x_head"]) and np.array_equal(y[:4], z["y_head"])        # same synthetic data as the recording
    layer = tnb.TensorTrainLayer(3, R, F + 1, output_shape=1, constrict_bond=True, perturb=True, seed=42)
    layer.to(device)
    tn = layer.tensor_network
    tn.gram_mode = gram_mode
    tn.small_site_fp64 = 0          # exercise the requested Gram mode even where the site is small
    trace = []
    ok = tn.accumulating_swipe(torch.tensor(X, device=device), torch.tensor(y, device=device), tnb.SquareBregFunction(), batch_size=512, lr=1.0,
                               eps=EPSS, orthonormalize=False, method="ridge_cholesky", num_swipes=NUM_SWIPES, skip_second=False, direction="l2r",
                               loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    ref = z["trace"]
    assert ok == bool(z["ok"])
    assert [(a, b) for a, b, _ in trace] == [(int(a), int(b)) for a, b, _ in ref]
    loss_err = np.array([abs(t[2] - r[2]) / max(abs(r[2]), 1e-300) for t, r in zip(trace, ref)])
    pred = tn.forward(torch.tensor(X[:256], device=device), to_tensor=True).cpu().numpy()
    pred_err = gu.relerr(pred.reshape(z["pred256"].shape), z["pred256"])
    core_err = max(gu.relerr(nd.tensor.cpu().numpy(), z[f"core_{i}"]) for i, nd in enumerate(tn.train_nodes))
    return loss_err, pred_err, core_err
