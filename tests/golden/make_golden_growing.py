"""Golden recording of the reference's growing-TT estimator (TensorTrainRegressorEarlyStopping, tensor/module.py:502-614)
-- build container only.   python tests/golden/make_golden_growing.py"""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.module import TensorTrainRegressorEarlyStopping  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(11)
    N, F = 400, 4
    X = rng.uniform(-1, 1, size=(N, F))
    y = 0.5 * X[:, 0] - X[:, 1] * X[:, 2] + 0.7 * X[:, 0] * X[:, 1] * X[:, 3] + 0.02 * rng.normal(size=N)
    Xv = rng.uniform(-1, 1, size=(150, F))
    yv = 0.5 * Xv[:, 0] - Xv[:, 1] * Xv[:, 2] + 0.7 * Xv[:, 0] * Xv[:, 1] * Xv[:, 3]
    flat = {"X": X, "y": y, "Xv": Xv, "yv": yv}
    for tag, kw in (("a", dict(N=6, r=4, eps_start=1e-2, eps_end=1e-6, early_stopping=2)),
                    ("b", dict(N=5, r=3, eps_start=1.0, eps_end=1e-3, early_stopping=10, constrict_bond=False))):
        est = TensorTrainRegressorEarlyStopping(device="cpu", batch_size=128, seed=3, **kw)
        est.fit(X, y, X_val=Xv, y_val=yv)
        hist = est._early_stopping.val_history
        flat[f"{tag}_pred"] = est.predict(Xv)
        flat[f"{tag}_best_degree"] = np.array(est._best_degree)
        flat[f"{tag}_singular"] = np.array(est._singular)
        flat[f"{tag}_val_history"] = np.array([hist[k] for k in sorted(hist)])
        flat[f"{tag}_score"] = np.array(est.score(Xv, yv))
        for i, nd in enumerate(est._model.tensor_network.train_nodes):
            flat[f"{tag}_core_{i}"] = nd.tensor.detach().numpy()
        print(tag, "best degree", est._best_degree, "history", flat[f"{tag}_val_history"], "score", flat[f"{tag}_score"])
    np.savez_compressed(os.path.join(OUT, "growing_tt.npz"), **flat)


if __name__ == "__main__":
    main()
