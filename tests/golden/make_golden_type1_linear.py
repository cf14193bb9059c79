"""Golden recording of the reference's type-I model built from TensorTrainLinearLayer members (models/tensor_train.py:177-188),
train_operators=True so the projections are trained too -- build container only."""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLinearLayer, TensorNetworkLayer  # noqa: E402
from tensor.network import SumOfNetworks  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(34)
    N, F, r, NN, lin, seed = 260, 5, 3, 3, 3, 42
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1] + X[:, 3:4]) + X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1)))
    f = F + 1
    nets = [TensorTrainLinearLayer(i, bond_dim=r, input_features=f - 1 if i != 1 else f, linear_dim=lin, output_shape=1, constrict_bond=False,
                                   perturb=False, seed=seed + i).tensor_network for i in range(1, NN + 1)]
    model = TensorNetworkLayer(SumOfNetworks(nets, output_labels=nets[0].output_labels, train_operators=True))
    tn = model.tensor_network
    flat = {"x": Xb.numpy(), "y": y.numpy(), "n_cores": np.array(len(tn.train_nodes)), "names": np.array([n.name for n in tn.train_nodes])}
    for i, nd in enumerate(tn.train_nodes):
        flat[f"cores0_{i}"] = nd.tensor.detach().numpy().copy()
    flat["pred0"] = tn.forward(Xb, to_tensor=True).detach().numpy().copy()
    trace = []
    tn.accumulating_swipe(Xb, y, SquareBregFunction(), batch_size=100, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.5, eps_decay=0.5,
                          loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    flat["trace"] = np.array(trace)
    for i, nd in enumerate(tn.train_nodes):
        flat[f"final_{i}"] = nd.tensor.detach().numpy().copy()
    flat["pred"] = tn.forward(Xb, to_tensor=True).detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "type1_linear.npz"), **flat)
    print("type1_linear:", list(flat["names"]), len(trace), "updates", [round(t[2], 6) for t in trace][:6])


if __name__ == "__main__":
    main()
