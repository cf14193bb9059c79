"""Recordings of the reference's accumulating_swipe(method='gradient') (tensor/network.py:458-470: every minibatch applies
theta += lr * J^T g at the current theta; the accumulated A is a random dummy and the solve is skipped) -- build container only.

    python tests/golden/make_golden_gradient.py

Stored in the layout tests/golden_util.load_krylov reads: cores before, all cores after every node update, per-node losses.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (stubs matplotlib, puts /root/reference on sys.path)
import torch  # noqa: E402
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction, XEAutogradBregman  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def record(name, layer, X, y, loss_fn, **kw):
    tn = layer.tensor_network
    cores0 = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    ups, losses = [], []

    def block_callback(NS, node):
        ups.append({"NS": NS, "k": tn.train_nodes.index(node), "after": [n.tensor.detach().numpy().copy() for n in tn.train_nodes]})

    ok = tn.accumulating_swipe(X, y, loss_fn, method="gradient", block_callback=block_callback,
                               loss_callback=lambda NS, node, l: losses.append(float(l)), **kw)
    assert ok
    flat = {"x": X.numpy(), "y": y.numpy(), "n_cores": np.array(len(cores0)), "n_updates": np.array(len(ups)), "losses": np.array(losses)}
    for i, c in enumerate(cores0):
        flat[f"cores0_{i}"] = c
    for ui, u in enumerate(ups):
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"]])
        for i, c in enumerate(u["after"]):
            flat[f"u{ui}_after_{i}"] = c
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print(name, len(ups), "updates; losses", losses[0], "->", losses[-1])


def main():
    rng = np.random.default_rng(7)
    N, F = 250, 3
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    # regression, minibatches of 64 (the last one short), descent through a negative lr, step control on
    y = torch.tensor(mg.teacher(X, 5))
    layer = TensorTrainLayer(3, 3, F + 1, output_shape=1, constrict_bond=False, perturb=True, seed=31)
    record("grad_tt_reg", layer, Xb, y, SquareBregFunction(), batch_size=64, num_swipes=2, lr=-2e-3, adaptive_step=True, max_norm=5.0)
    # classifier: class leg on the first core, cross-entropy, one batch
    lab = np.argmax(X @ rng.normal(size=(F, 3)), axis=1)
    y = torch.tensor(np.eye(3)[lab])
    layer = TensorTrainLayer(3, 3, F + 1, output_shape=2, constrict_bond=False, perturb=False, seed=32)
    record("grad_tt_xe", layer, Xb, y, XEAutogradBregman(w=1.0), batch_size=-1, num_swipes=1, lr=-1e-3)
    record_type1(Xb, torch.tensor(mg.teacher(X, 6)))


def record_type1(Xb, y):
    """Sum of trains with 1..3 cores (the tt_type1 models): the member's minibatch step sees the other members' outputs."""
    from tensor.network import SumOfNetworks
    from tensor.layers import TensorNetworkLayer
    f = Xb.shape[1]
    nets = [TensorTrainLayer(i, bond_dim=2, input_features=f - 1 if i != 1 else f, output_shape=1, constrict_bond=True, perturb=True,
                             seed=40 + i).tensor_network for i in range(1, 4)]
    layer = TensorNetworkLayer(SumOfNetworks(nets, output_labels=nets[0].output_labels))
    record("grad_type1", layer, Xb, y, SquareBregFunction(), batch_size=100, num_swipes=1, lr=-1e-3)


if __name__ == "__main__":
    main()
