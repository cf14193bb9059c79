"""BASELINE config 4 in its conv-TT reading (4b) at full model size on a row subsample, recorded from the unmodified reference
(build container only):

    python tests/golden/make_golden_cfg4b.py

MNIST-shaped patches: x (N, 50 patches, 17 pixels) with the bias patch / bias pixel of image_convolution_CG_MNIST.py:29-32,
TensorConvolutionTrainLayer(num_carriages=3, bond_dim=38, num_patches=50, patch_pixels=17, output_shape=9, convolution_bond=4)
(patch cores of 17 100, 72 200 and 1 900 parameters), cross-entropy with w = 1, scipy_swipe(minres) as image_convolution_CG_MNIST.py:95
runs it, here with 3 Krylov steps per node on 256 rows (the reference materialises the batch Jacobian of the 72 200-parameter core
for every matvec: 1.3 GB at 256 rows).  The tests regenerate the data.
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
from scipy.sparse.linalg import minres  # noqa: E402
from tensor.layers import TensorConvolutionTrainLayer  # noqa: E402
from tensor.bregman import XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N, Q, T, R, CB, C = 256, 50, 17, 38, 4, 9


def data():
    rng = np.random.default_rng(2028)
    X = rng.uniform(0, 1, size=(N, Q, T))
    X[:, -1, :] = 0.0
    X[:, :, -1] = 0.0
    X[:, -1, -1] = 1.0
    y = np.eye(C + 1)[rng.integers(0, C + 1, N)]
    return X, y


def main():
    X, y = data()
    torch.manual_seed(42)
    layer = TensorConvolutionTrainLayer(num_carriages=3, bond_dim=R, num_patches=Q, patch_pixels=T, output_shape=C, convolution_bond=CB)
    tn = layer.tensor_network
    core_sums = np.array([float(n.tensor.double().abs().sum()) for n in tn.train_nodes])     # the mirrored constructor must draw the same cores
    losses = []
    ok = tn.scipy_swipe(torch.tensor(X), torch.tensor(y), XEAutogradBregman(w=1.0), minres, batch_size=256, num_swipes=1, lr=1.0, max_iter=3,
                        tol=1e-3, loss_callback=lambda l: (losses.append(float(l)), print(len(losses), float(l), flush=True)))
    tn.reset_stacks()
    pred = tn.forward(torch.tensor(X[:64]), to_tensor=True).detach().numpy()
    flat = dict(ok=np.array(bool(ok)), losses=np.array(losses), pred64=pred, x_head=X[:2], names=np.array([n.name for n in tn.train_nodes]),
                core_abs_sums0=core_sums)
    np.savez_compressed(os.path.join(OUT, "cfg4b_shape.npz"), **flat)
    print("ok", ok, losses)


if __name__ == "__main__":
    main()
