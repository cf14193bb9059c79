"""Shared body of the method='gradient' parity tests: tests/golden/make_golden_gradient.py."""
import numpy as np
import torch

import golden_util as gu
import tensornetworksfork_b200 as tnb


def _tt(n, r, f, C, perturb, seed):
    return tnb.TensorTrainLayer(n, r, f, output_shape=C, constrict_bond=False, perturb=perturb, seed=seed)


def _type1(f):
    nets = [tnb.TensorTrainLayer(i, bond_dim=2, input_features=f - 1 if i != 1 else f, output_shape=1, constrict_bond=True, perturb=True,
                                 seed=40 + i).tensor_network for i in range(1, 4)]
    return tnb.TensorNetworkLayer(tnb.SumOfNetworks(nets, output_labels=nets[0].output_labels))


CASES = {
    "grad_tt_reg": dict(build=lambda: _tt(3, 3, 4, 1, True, 31), loss=lambda: tnb.SquareBregFunction(),
                        kw=dict(batch_size=64, num_swipes=2, lr=-2e-3, adaptive_step=True, max_norm=5.0)),
    "grad_tt_xe": dict(build=lambda: _tt(3, 3, 4, 2, False, 32), loss=lambda: tnb.XEAutogradBregman(w=1.0),
                       kw=dict(batch_size=-1, num_swipes=1, lr=-1e-3)),
    "grad_type1": dict(build=lambda: _type1(4), loss=lambda: tnb.SquareBregFunction(), kw=dict(batch_size=100, num_swipes=1, lr=-1e-3)),
}


def run(name, device, group=None, shard=None):
    """(max relative core error over all updates, max loss error).  ``shard = (lo, hi)`` binds that row range under ``group``."""
    case, fx = CASES[name], gu.load_krylov(name)
    layer = case["build"]()
    tn = layer.tensor_network
    for n, c in zip(tn.train_nodes, fx["cores0"]):                 # same seeds -> same initial cores as the reference
        assert np.array_equal(n.tensor.numpy(), c), n.name
    layer.to(device)
    X, y = torch.tensor(fx["x"], device=device), torch.tensor(fx["y"], device=device)
    if shard is not None:
        tn.process_group, tn.shard_offset, tn.shard_total = group, shard[0], X.shape[0]
        X, y = X[shard[0]:shard[1]].contiguous(), y[shard[0]:shard[1]].contiguous()
    ups, losses = [], []
    ok = tn.accumulating_swipe(X, y, case["loss"](), method="gradient", loss_callback=lambda NS, node, l: losses.append(l),
                               block_callback=lambda NS, node: ups.append((NS, tn.train_nodes.index(node),
                                                                           [n.tensor.cpu().numpy().copy() for n in tn.train_nodes])),
                               **case["kw"])
    assert ok
    assert [(a, b) for a, b, _ in ups] == [(u["NS"], u["k"]) for u in fx["updates"]]
    core_err = max(gu.relerr(c, ref) for (_, _, cores), u in zip(ups, fx["updates"]) for c, ref in zip(cores, u["after"]))
    loss_err = float(np.max(np.abs(np.array(losses) - fx["losses"]) / np.maximum(1.0, np.abs(fx["losses"]))))
    return core_err, loss_err
