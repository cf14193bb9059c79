"""QR re-gauge of wide cores (fewer rows than columns): the reference's ``torch.linalg.qr(mode='reduced')`` returns an m x m Q there, so
the bond SHRINKS to the core's row count (reference network.py:644-657, 686-704) -- an unconstricted train with r > f, e.g.
``TNMLRegressor(constrict_bond=False)``.  Shared body of the GPU test and its CPU twin: a left sweep of re-gauges, then a right sweep,
against the numpy oracle; the network function must not change."""
import numpy as np
import torch


def run(device, sites=4, r=5, f=2, seed=3):
    import tensornetworksfork_b200 as tnb
    from oracle import tn_oracle as orc
    layer = tnb.TensorTrainLayer(sites, r, f, output_shape=1, constrict_bond=False, seed=seed)
    tn = layer.tensor_network
    cores = [n.tensor.numpy().copy() for n in tn.main_nodes]
    X = np.random.default_rng(seed).uniform(-1, 1, size=(64, f))
    layer.to(device)
    x = torch.tensor(X, device=device)
    p0 = tn.forward(x, to_tensor=True).cpu().numpy()
    for k in range(len(cores) - 1):
        tn.node_orthonormalize_left(tn.main_nodes[k])
        orc.orthonormalize_left(cores, k)
    mid = [tuple(n.tensor.shape) for n in tn.main_nodes]
    for k in range(len(cores) - 1, 0, -1):
        tn.node_orthonormalize_right(tn.main_nodes[k])
        orc.orthonormalize_right(cores, k)
    worst = 0.0
    for n, c in zip(tn.main_nodes, cores):
        g = n.tensor.cpu().numpy()
        assert g.shape == c.shape, (g.shape, c.shape)
        worst = max(worst, float(np.linalg.norm(g - c) / np.linalg.norm(c)))
    p1 = tn.forward(x, to_tensor=True).cpu().numpy()
    drift = float(np.linalg.norm(p1 - p0) / np.linalg.norm(p0))
    return mid, [tuple(c.shape) for c in cores], worst, drift
