"""Drop-in check (build container only: needs /root/reference): the B200 engine class accepts the node graph built by
the REFERENCE's own layer constructors, and -- driven on the CPU stand-in kernels -- reproduces the reference's sweep.
Skipped where the reference tree is absent (the GPU box)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tensor")), reason="reference tree not mounted")

torch.set_default_dtype(torch.float64)


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)          # tensor/utils.py:2 imports pyplot; plotting unused
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.append(REF)
    import importlib
    ref_layers = importlib.import_module("tensor.layers")
    ref_breg = importlib.import_module("tensor.bregman")
    return ref_layers, ref_breg


def test_engine_runs_reference_built_graph(monkeypatch):
    import fake_ops
    ref_layers, ref_breg = _import_reference()
    from tensornetworksfork_b200.tensor.network import TensorNetwork as FastTN
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(0)
    N, F = 200, 4
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1].numpy()) + 0.1 * rng.normal(size=(N, 1)))
    kw = dict(batch_size=64, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)

    ref_layer = ref_layers.TensorTrainLayer(4, 3, F + 1, output_shape=1, constrict_bond=False, seed=5)
    ref_losses = []
    ref_layer.tensor_network.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: ref_losses.append(l), **kw)
    ref_pred = ref_layer.tensor_network.forward_batch(X, 64)

    layer2 = ref_layers.TensorTrainLayer(4, 3, F + 1, output_shape=1, constrict_bond=False, seed=5)
    old = layer2.tensor_network
    fast = FastTN(old.input_nodes, old.main_nodes, old.train_nodes, output_labels=old.output_labels, sample_dim=old.sample_dim)
    layer2.set_tensor_network(fast)                     # the reference's own layer object now owns the B200 engine
    losses = []
    ok = fast.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: losses.append(l), **kw)
    assert ok
    assert len(losses) == len(ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b))
    pred = fast.forward_batch(X, 64)
    assert float((pred - ref_pred).norm() / ref_pred.norm()) < 1e-8
    states = layer2.node_states()                       # reference's checkpoint hooks still work on the new engine
    layer2.load_node_states(states, set_value=True)
    assert float((fast.forward_batch(X, 64) - ref_pred).norm() / ref_pred.norm()) < 1e-8
