"""Golden recording of the reference's type-I model (SumOfNetworks of tensor trains with 1..N cores) -- build container only.
Construction follows models/tensor_train.py:165-176."""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLayer, TensorNetworkLayer  # noqa: E402
from tensor.network import SumOfNetworks  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(33)
    N, F, r, NN, seed = 240, 3, 3, 3, 42
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1]) + X[:, 1:2] * X[:, 2:3] + 0.05 * rng.normal(size=(N, 1)))
    f = F + 1
    nets = [TensorTrainLayer(i, bond_dim=r, input_features=f - 1 if i != 1 else f, output_shape=1, constrict_bond=False, perturb=False,
                             seed=seed + i).tensor_network for i in range(1, NN + 1)]
    model = TensorNetworkLayer(SumOfNetworks(nets, output_labels=nets[0].output_labels, train_operators=False))
    tn = model.tensor_network
    flat = {"x": Xb.numpy(), "y": y.numpy(), "n_cores": np.array(len(tn.train_nodes))}
    for i, nd in enumerate(tn.train_nodes):
        flat[f"cores0_{i}"] = nd.tensor.detach().numpy().copy()
    flat["pred0"] = tn.forward(Xb, to_tensor=True).detach().numpy().copy()
    trace = []
    tn.accumulating_swipe(Xb, y, SquareBregFunction(), batch_size=80, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5, eps_decay=0.5,
                          loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    flat["trace"] = np.array(trace)
    for i, nd in enumerate(tn.train_nodes):
        flat[f"final_{i}"] = nd.tensor.detach().numpy().copy()
    flat["pred"] = tn.forward(Xb, to_tensor=True).detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "type1_tt.npz"), **flat)
    print("type1_tt:", len(trace), "updates", [round(t[2], 6) for t in trace][:8])


if __name__ == "__main__":
    main()
