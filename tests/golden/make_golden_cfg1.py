"""BASELINE config 1 at FULL size, recorded from the unmodified reference (build container only):

    python tests/golden/make_golden_cfg1.py        # ~1 minute

TT regression on synthetic abalone-shaped data (N = 4177, 8 features + bias, 3 cores, rank 6), the call of default_train.py:101-129
-- TensorTrainLayer(3, r, F + 1, output_shape=1, constrict_bond=True, perturb=True, seed=42), batch_size=512, ridge_cholesky,
eps = geomspace(0.0754, 7.2e-12, 8), NUM_SWIPES = 4 -- with r = 6 as BASELINE.json substitutes.  The data are regenerated from
the seed by the tests; the fixture keeps the per-update (NS, node, loss) trace, the final prediction on the first 256 rows and the
final cores.
"""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N, F, R, NUM_SWIPES = 4177, 8, 6, 4
EPSS = np.geomspace(0.07542717629430484, 0.00000000000722857583, 2 * NUM_SWIPES).tolist()


def data():
    """U(-1, 1) features + bias column; a smooth degree-3 teacher with noise."""
    rng = np.random.default_rng(2024)
    X = rng.uniform(-1, 1, size=(N, F))
    W1, W2 = rng.normal(size=(F, 1)) / np.sqrt(F), rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W1) + 0.5 * (X @ W2) ** 2 + 0.3 * X[:, :1] * X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1))
    return np.concatenate([X, np.ones((N, 1))], 1), y


def main():
    X, y = data()
    layer = TensorTrainLayer(3, R, F + 1, output_shape=1, constrict_bond=True, perturb=True, seed=42)
    tn = layer.tensor_network
    trace = []
    ok = tn.accumulating_swipe(torch.tensor(X), torch.tensor(y), SquareBregFunction(), batch_size=512, lr=1.0, eps=EPSS, orthonormalize=False,
                               method="ridge_cholesky", num_swipes=NUM_SWIPES, skip_second=False, direction="l2r",
                               loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    pred = tn.forward(torch.tensor(X[:256]), to_tensor=True).detach().numpy()
    flat = {"ok": np.array(bool(ok)), "trace": np.array(trace), "pred256": pred, "x_head": X[:4], "y_head": y[:4]}
    for i, nd in enumerate(tn.train_nodes):
        flat[f"core_{i}"] = nd.tensor.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "cfg1_full.npz"), **flat)
    print("ok", ok, len(trace), "updates; loss", trace[0][2], "->", trace[-1][2])


if __name__ == "__main__":
    main()
