"""Golden recording of the reference's growing conv-TT flow (TensorConvolutionTrainLayer.grow_cart, tensor/layers.py:892-947;
call shape of image_convolution_growing_MNIST.py:84-103) -- build container only.

    python tests/golden/make_golden_conv_grow.py

Three phases on one layer: a dense sweep on 2 columns, grow_cart(r, CB) + a sweep in direction 'r2l', grow_cart() with the default
bonds + a minibatched sweep.  Every phase stores the train-node names, the cores it starts from (after the growth), the prediction
before and after, and the cores after every node update.  ``seed_*`` holds the cores of a fresh layer grown twice without any
sweep, for the bit-for-bit constructor check (the new pixel cores are random draws).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_conv as mg  # noqa: E402  (stubs matplotlib, puts the reference on sys.path)
import torch  # noqa: E402
from tensor.layers import TensorConvolutionTrainLayer  # noqa: E402
from tensor.bregman import XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CTOR = dict(num_carriages=2, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2)
PHASES = [
    dict(grow=None, kw=dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_exact", eps=[1.0, 0.5])),
    dict(grow=(3, 2), kw=dict(batch_size=-1, num_swipes=1, lr=1.0, method="ridge_exact", eps=[0.8, 0.4], direction="r2l")),
    dict(grow=(None, None), kw=dict(batch_size=48, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.6, eps_decay=0.5)),
]


def main():
    flat = {}
    torch.manual_seed(33)
    fresh = TensorConvolutionTrainLayer(**CTOR)
    fresh.grow_cart(3, 2)
    fresh.grow_cart()
    names = [n.name for n in fresh.tensor_network.train_nodes]
    flat["seed_names"] = np.array(names)
    for i, n in enumerate(fresh.tensor_network.train_nodes):
        flat[f"seed_core_{i}"] = n.tensor.detach().numpy().copy()

    torch.manual_seed(33)
    X, y = mg.data(9, 110, 5, 4, K=3)
    layer = TensorConvolutionTrainLayer(**CTOR)
    loss_fn = XEAutogradBregman(w=1.0)
    for pi, ph in enumerate(PHASES):
        if ph["grow"] is not None:
            layer.grow_cart(*ph["grow"])
        tn = layer.tensor_network
        names = [n.name for n in tn.train_nodes]
        cores0 = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
        pred0 = tn.forward(X, to_tensor=True).detach().numpy().copy()
        tn.reset_stacks()
        ups, losses = [], []

        def block_callback(NS, node, tn=tn, ups=ups):
            ups.append({"NS": NS, "k": tn.train_nodes.index(node), "after": [n.tensor.detach().numpy().copy() for n in tn.train_nodes]})

        ok = tn.accumulating_swipe(X, y, loss_fn, block_callback=block_callback,
                                   loss_callback=lambda NS, node, l, losses=losses: losses.append(float(l)), **ph["kw"])
        assert ok
        tn.reset_stacks()
        pred1 = tn.forward(X, to_tensor=True).detach().numpy().copy()
        pre = f"p{pi}_"
        flat.update({pre + "names": np.array(names), pre + "n_cores": np.array(len(cores0)), pre + "n_updates": np.array(len(ups)),
                     pre + "losses": np.array(losses), pre + "pred0": pred0, pre + "pred1": pred1})
        for i, c in enumerate(cores0):
            flat[pre + f"cores0_{i}"] = c
        for ui, u in enumerate(ups):
            flat[pre + f"u{ui}_scal"] = np.array([u["NS"], u["k"]])
            for i, c in enumerate(u["after"]):
                flat[pre + f"u{ui}_after_{i}"] = c
        print("phase", pi, names, len(ups), "updates; losses", losses[0], "->", losses[-1])
    flat["x"], flat["y"] = X.numpy(), y.numpy()
    np.savez_compressed(os.path.join(OUT, "conv_grow.npz"), **flat)


if __name__ == "__main__":
    main()
