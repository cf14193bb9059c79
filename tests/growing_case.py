"""Shared body of the growing-TT (one-pass, per-core epsilon, early stopping) parity test."""
import os

import numpy as np

import golden_util as gu

KW = {"a": dict(N=6, r=4, eps_start=1e-2, eps_end=1e-6, early_stopping=2),
      "b": dict(N=5, r=3, eps_start=1.0, eps_end=1e-3, early_stopping=10, constrict_bond=False)}


def run(tag, device):
    from tensornetworksfork_b200.tensor.module import TensorTrainRegressorEarlyStopping
    z = np.load(os.path.join(gu.GOLDEN_DIR, "growing_tt.npz"))
    est = TensorTrainRegressorEarlyStopping(device=device, batch_size=128, seed=3, **KW[tag])
    est.fit(z["X"], z["y"], X_val=z["Xv"], y_val=z["yv"])
    hist = est._early_stopping.val_history
    got_hist = np.array([hist[k] for k in sorted(hist)])
    assert est._best_degree == int(z[f"{tag}_best_degree"])
    assert est._singular == bool(z[f"{tag}_singular"])
    assert got_hist.shape == z[f"{tag}_val_history"].shape
    hist_err = float(np.max(np.abs(got_hist - z[f"{tag}_val_history"])))
    pred_err = gu.relerr(est.predict(z["Xv"]), z[f"{tag}_pred"])
    core_err = max(gu.relerr(nd.tensor.cpu().numpy(), z[f"{tag}_core_{i}"]) for i, nd in enumerate(est._model.tensor_network.train_nodes))
    score_err = abs(est.score(z["Xv"], z["yv"]) - float(z[f"{tag}_score"]))
    return hist_err, pred_err, core_err, score_err
