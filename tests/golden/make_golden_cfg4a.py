"""BASELINE config 4 in its TNML reading (4a) at full local size on a short chain, recorded from the unmodified reference (build
container only):

    python tests/golden/make_golden_cfg4a.py

TNML classifier with the sin-cos map, 10 classes -> 9 logits on the first core, cross-entropy (XEAutogradBregman, w = 1), rank 38
with constricted bonds (1, 2, 4, 8, 16, 32, 38, ..., 38, 32, ..., 1: local systems up to 38 * 2 * 38 = 2888 parameters, the
largest of the 784-site configuration), local solves by scipy_swipe(cg) as image_convolution_CG_MNIST.py:95 calls it (float32
Krylov recurrences on the host).  16 sites (a 4 x 4 image) instead of 784 -- the reference rebuilds all environments for every
matvec, a pass over 784 sites takes it the better part of an hour -- on 512 rows.  The tests regenerate the data.
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
from scipy.sparse.linalg import cg  # noqa: E402
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N, SITES, R, C = 512, 16, 38, 9


def data():
    rng = np.random.default_rng(2030)
    X = rng.uniform(0, 1, size=(N, SITES))
    y = np.eye(C + 1)[np.argmax(X @ rng.normal(size=(SITES, C + 1)), axis=1)]
    return X, y


def main():
    X, y = data()
    xs = [torch.tensor(np.stack([np.cos(0.5 * np.pi * X[:, j]), np.sin(0.5 * np.pi * X[:, j])], 1)) for j in range(SITES)]
    layer = TensorTrainLayer(SITES, R, 2, output_shape=C, constrict_bond=True, seed=42)
    tn = layer.tensor_network
    losses = []
    ok = tn.scipy_swipe(xs, torch.tensor(y), XEAutogradBregman(w=1.0), cg, batch_size=512, num_swipes=2, lr=0.05, max_iter=5, tol=1e-3,
                        loss_callback=lambda l: losses.append(float(l)))
    tn.reset_stacks()
    pred = tn.forward([t[:128] for t in xs], to_tensor=True).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "cfg4a_chain16.npz"), ok=np.array(bool(ok)), losses=np.array(losses), pred128=pred, x_head=X[:2],
                        shapes=np.array([list(n.tensor.shape) + [0] * (3 - n.tensor.dim()) for n in tn.train_nodes]))
    print("ok", ok, len(losses), "updates; loss", losses[0], "->", losses[-1], "max P", max(n.tensor.numel() for n in tn.train_nodes))


if __name__ == "__main__":
    main()
