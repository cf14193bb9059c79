"""Type-I model (SumOfNetworks) whose members are linear-projection trains, projections trained (reference
models/tensor_train.py:177-188): host logic on the CPU stand-in kernels against a recording of the reference."""
import os

import numpy as np
import pytest
import torch

import fake_ops
import golden_util as gu
import tensornetworksfork_b200 as tnb

torch.set_default_dtype(torch.float64)


def run(device):
    z = np.load(os.path.join(gu.GOLDEN_DIR, "type1_linear.npz"))
    X, y = torch.tensor(z["x"], device=device), torch.tensor(z["y"], device=device)
    f, r, NN, lin, seed = 6, 3, 3, 3, 42
    nets = [tnb.TensorTrainLinearLayer(i, bond_dim=r, input_features=f - 1 if i != 1 else f, linear_dim=lin, output_shape=1,
                                       constrict_bond=False, perturb=False, seed=seed + i).tensor_network for i in range(1, NN + 1)]
    model = tnb.TensorNetworkLayer(tnb.SumOfNetworks(nets, output_labels=nets[0].output_labels, train_operators=True))
    tn = model.tensor_network
    assert len(tn.train_nodes) == int(z["n_cores"])
    for i, nd in enumerate(tn.train_nodes):
        assert np.array_equal(nd.tensor.numpy(), z[f"cores0_{i}"])
    model.to(device)
    assert gu.relerr(tn.forward(X, to_tensor=True).cpu().numpy().reshape(z["pred0"].shape), z["pred0"]) < 1e-12
    trace = []
    assert tn.accumulating_swipe(X, y, tnb.SquareBregFunction(), batch_size=100, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=0.5,
                                 eps_decay=0.5, loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), l)))
    ref = z["trace"]
    assert [(a, b) for a, b, _ in trace] == [(int(a), int(b)) for a, b, _ in ref]
    for (_, _, l), (_, _, rl) in zip(trace, ref):
        assert abs(l - rl) <= 1e-7 * max(1.0, abs(rl))
    assert gu.relerr(tn.forward(X, to_tensor=True).cpu().numpy().reshape(z["pred"].shape), z["pred"]) < 1e-7
    for i, nd in enumerate(tn.train_nodes):
        assert gu.relerr(nd.tensor.cpu().numpy(), z[f"final_{i}"]) < 1e-6


def test_type1_of_linear_trains(monkeypatch):
    fake_ops.install(monkeypatch)
    run("cpu")


@pytest.mark.gpu
def test_type1_of_linear_trains_gpu():
    run("cuda")
