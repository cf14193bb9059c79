"""BASELINE config 5 in its primary reading (5a: TT poly-mode, 5 cores, rank 38, 28 features + bias, local systems of P = 41 876) --
a fingerprint of the reference's own get_A_b at that shape (build container only; needs ~30 GB of RAM):

    python tests/golden/make_golden_cfg5a.py

A whole site update of the reference at this shape needs ~56 GB (A is 14 GB and solve_system copies it three times), so the
recording stops after get_A_b (network.py:174-217) on ONE 512-row minibatch at the middle core: the prediction, the per-row loss,
b (P), diag(A) (P) and A v for a seeded random v (P) -- enough to pin every entry class of the 41 876 x 41 876 Gram matrix.
opt_einsum stand-in as in make_golden_cfg2.py.  The tests regenerate the data and the cores (same seed).
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
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

N, F, R, CORES, SITE = 512, 28, 38, 5, 2


def data():
    rng = np.random.default_rng(2029)
    X = rng.uniform(-1, 1, size=(N, F))
    W = rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1))
    return np.concatenate([X, np.ones((N, 1))], 1), y


def main():
    assert torch.backends.opt_einsum.is_available()
    X, y = data()
    layer = TensorTrainLayer(CORES, R, F + 1, output_shape=1, constrict_bond=False, seed=42)
    tn = layer.tensor_network
    node = tn.train_nodes[SITE]
    P = node.tensor.numel()
    with torch.no_grad():
        pred = tn.forward(torch.tensor(X), to_tensor=True)
        loss, g, H = SquareBregFunction().forward(pred, torch.tensor(y))
        A, b = tn.get_A_b(node, g, H)
        A = A.reshape(P, P)
        v = torch.tensor(np.random.default_rng(7).normal(size=P))
        Av = A @ v
        flat = dict(pred=pred.numpy(), loss=loss.numpy(), b=b.reshape(P).numpy(), diagA=A.diagonal().numpy().copy(), Av=Av.numpy(),
                    fro=np.array(float(torch.linalg.matrix_norm(A))), asym=np.array(float((A - A.t()).abs().max())),
                    x_head=X[:2], core_abs_sums=np.array([float(n.tensor.abs().sum()) for n in tn.train_nodes]))
    np.savez_compressed(os.path.join(OUT, "cfg5a_gram.npz"), **flat)
    print("P", P, "fro", flat["fro"], "asym", flat["asym"], "|b|", np.linalg.norm(flat["b"]))


if __name__ == "__main__":
    main()
