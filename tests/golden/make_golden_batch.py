"""Golden recording of the reference's minibatch estimator (TensorTrainBatchRegressor, tensor/module.py:308-500; call shape of
train_mnist_batch.py:54-73) in its three swipe methods -- build container only.   python tests/golden/make_golden_batch.py"""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.module import TensorTrainBatchRegressor  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"unique": dict(swipe_method="batch_unique", num_swipes=3, N=3, r=3, batch_size=64, eps_start=2.0, eps_end=0.5),
         "same": dict(swipe_method="batch_same", num_swipes=2, N=3, r=3, batch_size=96, eps_start=2.0, eps_end=0.5, perturb=False),
         "block": dict(swipe_method="batch_block", num_swipes=2, N=4, r=2, batch_size=80, eps_start=1.0, eps_end=1.0)}


def main():
    rng = np.random.default_rng(17)
    N, F = 330, 4
    X = rng.uniform(-1, 1, size=(N, F))
    y = 0.5 * X[:, 0] - X[:, 1] * X[:, 2] + 0.7 * X[:, 0] * X[:, 1] * X[:, 3] + 0.02 * rng.normal(size=N)
    flat = {"X": X, "y": y}
    for tag, kw in CASES.items():
        est = TensorTrainBatchRegressor(device="cpu", seed=5, **kw)
        est.fit(X, y)                     # validation split 0.1 drawn from the seed
        flat[f"{tag}_pred"] = est.predict(X)
        flat[f"{tag}_traj"] = np.array([[t["epoch"], t["val_rmse"]] for t in est.trajectory])
        for i, nd in enumerate(est._model.tensor_network.train_nodes):
            flat[f"{tag}_core_{i}"] = nd.tensor.detach().numpy()
        print(tag, "trajectory", flat[f"{tag}_traj"].tolist())
    np.savez_compressed(os.path.join(OUT, "batch_tt.npz"), **flat)


if __name__ == "__main__":
    main()
