"""Shared body of the type-I (SumOfNetworks) parity test."""
import os

import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb


def run(device):
    z = np.load(os.path.join(gu.GOLDEN_DIR, "type1_tt.npz"))
    X = torch.tensor(z["x"], device=device)
    y = torch.tensor(z["y"], device=device)
    f, r, NN, seed = 4, 3, 3, 42
    nets = [tnb.TensorTrainLayer(i, bond_dim=r, input_features=f - 1 if i != 1 else f, output_shape=1, constrict_bond=False,
                                 perturb=False, seed=seed + i).tensor_network for i in range(1, NN + 1)]
    model = tnb.TensorNetworkLayer(tnb.SumOfNetworks(nets, output_labels=nets[0].output_labels, train_operators=False))
    tn = model.tensor_network
    assert len(tn.train_nodes) == int(z["n_cores"])
    for i, nd in enumerate(tn.train_nodes):                       # same seeds -> same initial members as the reference
        assert np.array_equal(nd.tensor.numpy(), z[f"cores0_{i}"])
    model.to(device)
    pred0 = tn.forward(X, to_tensor=True).cpu().numpy()
    assert gu.relerr(pred0.reshape(z["pred0"].shape), z["pred0"]) < 1e-12
    trace = []
    ok = tn.accumulating_swipe(X, y, tnb.SquareBregFunction(), batch_size=80, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5,
                               eps_decay=0.5, loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), l)))
    assert ok
    ref = z["trace"]
    assert len(trace) == len(ref)
    for (NS, k, l), (rNS, rk, rl) in zip(trace, ref):
        assert (NS, k) == (int(rNS), int(rk))
        assert abs(l - rl) <= 1e-7 * max(1.0, abs(rl)), (NS, k, l, rl)
    pred = tn.forward(X, to_tensor=True).cpu().numpy()
    assert gu.relerr(pred.reshape(z["pred"].shape), z["pred"]) < 1e-7
    for i, nd in enumerate(tn.train_nodes):
        assert gu.relerr(nd.tensor.cpu().numpy(), z[f"final_{i}"]) < 1e-6
