"""BASELINE config 2 at FULL size, recorded from the unmodified reference (build container only):

    python tests/golden/make_golden_cfg2.py        # a few minutes

CPD model of rank 100 on synthetic california_housing-shaped data (N = 20640, 8 features + bias, 5 factors = degree 5):
CPDLayer(5, 100, 9, output_shape=(1,)) as default_CPD_house.py:70 / models/tensor_train.py:150 build it, swept with the wrapper
defaults of models/tensor_train.py:91-104 (ridge_cholesky, eps_start 1.0, eps_decay 0.5, batch_size 512), two sweeps.  The
reference's Gram einsum (network.py:212) is contracted in its authors' (J H)-first order by putting the opt_einsum stand-in of
tools/ref_vs_port.py on sys.path before torch is imported -- same code path, same numbers, minutes instead of hours
(baseline B).  The tests regenerate the data from the seed.
"""
import os
import sys
import types

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "tools"))
import ref_vs_port  # noqa: E402

oe_dir = "/tmp/tn_opt_einsum_standin"
os.makedirs(os.path.join(oe_dir, "opt_einsum"), exist_ok=True)
open(os.path.join(oe_dir, "opt_einsum", "__init__.py"), "w").write(ref_vs_port.STANDIN)
sys.path.insert(0, oe_dir)
m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import CPDLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

N, F, RANK, FACTORS, NUM_SWIPES = 20640, 8, 100, 5, 2


def data():
    rng = np.random.default_rng(2025)
    X = rng.uniform(-1, 1, size=(N, F))
    W1, W2 = rng.normal(size=(F, 1)) / np.sqrt(F), rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W1) + 0.5 * (X @ W2) ** 2 + 0.3 * X[:, :1] * X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1))
    return np.concatenate([X, np.ones((N, 1))], 1), y


def main():
    assert torch.backends.opt_einsum.is_available()
    X, y = data()
    layer = CPDLayer(FACTORS, RANK, F + 1, output_shape=(1,), seed=42)
    tn = layer.tensor_network
    trace = []
    ok = tn.accumulating_swipe(torch.tensor(X), torch.tensor(y), SquareBregFunction(), batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5,
                               orthonormalize=False, method="ridge_cholesky", num_swipes=NUM_SWIPES, skip_second=False, direction="l2r",
                               loss_callback=lambda NS, nd, l: (trace.append((NS, tn.train_nodes.index(nd), float(l))), print(trace[-1], flush=True)))
    pred = tn.forward(torch.tensor(X[:256]), to_tensor=True).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "cfg2_full.npz"), ok=np.array(bool(ok)), trace=np.array(trace), pred256=pred, x_head=X[:4], y_head=y[:4])
    print("ok", ok, len(trace), "updates; loss", trace[0][2], "->", trace[-1][2])


if __name__ == "__main__":
    main()
