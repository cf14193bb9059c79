"""Golden recording of the reference's growing 2-site (DMRG-style) flow -- build container only.
Call shape of growing_DMRG.py:51-62: sweep, then repeatedly grow_middle / sweep the block / split_node."""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainDMRGInfiLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(7)
    N, F, r = 400, 3, 4
    X = np.sort(rng.uniform(-1, 1, size=(N, F)), axis=1)
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(np.prod(X + 0.3, axis=1, keepdims=True) + 0.5 * X[:, :1] ** 2)
    torch.manual_seed(11)
    layer = TensorTrainDMRGInfiLayer(r, F + 1, output_shape=1, constrict_bond=True)
    flat = {"x": Xb.numpy(), "y": y.numpy()}
    losses = []
    stage = [0]

    def snap(tag):
        for i, n in enumerate(layer.nodes):
            flat[f"s{stage[0]}_{tag}_core{i}"] = n.tensor.detach().numpy().copy()
        flat[f"s{stage[0]}_{tag}_pred"] = layer.tensor_network.forward(Xb, to_tensor=True).detach().numpy().copy()

    snap("init")
    kw = dict(batch_size=-1, lr=1.0, orthonormalize=False, method="ridge_cholesky", num_swipes=5, skip_second=False, direction="l2r",
              loss_callback=lambda NS, n, l: losses.append(l))
    layer.tensor_network.accumulating_swipe(Xb, y, SquareBregFunction(), eps=1.0, **kw)
    snap("swept")
    for carts in range(3, 6):
        stage[0] += 1
        layer.grow_middle()
        snap("grown")
        layer.tensor_network.accumulating_swipe(Xb, y, SquareBregFunction(), eps=0.05, **kw)
        snap("swept")
        node = layer.nodes[layer.num_carriages // 2]
        err = layer.split_node(node.dim_labels[:2], node.dim_labels[-2:], r, err=1e-6, is_last=carts == 5)
        flat[f"s{stage[0]}_split_err"] = np.array(float(err))
        snap("split")
    flat["losses"] = np.array(losses)
    flat["n_stages"] = np.array(stage[0] + 1)
    np.savez_compressed(os.path.join(OUT, "dmrg_growing.npz"), **flat)
    print("dmrg_growing:", len(losses), "updates; losses", losses)


if __name__ == "__main__":
    main()
