"""QR re-gauge of wide cores (bond shrink, reference network.py:644-657) on the device against the numpy oracle: an unconstricted
4-site train with r = 5 > f = 2, left sweep of re-gauges then right sweep; prints the worst core error and the prediction drift."""
import os, sys, time
t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
torch.set_default_dtype(torch.float64)
import tensornetworksfork_b200 as tnb
from oracle import tn_oracle as orc
layer = tnb.TensorTrainLayer(4, 5, 2, output_shape=1, constrict_bond=False, seed=3)
tn = layer.tensor_network
cores = [n.tensor.numpy().copy() for n in tn.main_nodes]
X = np.random.default_rng(0).uniform(-1, 1, size=(64, 2))
layer.to("cuda:0")
x = torch.tensor(X, device="cuda:0")
p0 = tn.forward(x, to_tensor=True).cpu().numpy()
worst = 0.0
for k in range(len(cores) - 1):
    tn.node_orthonormalize_left(tn.main_nodes[k])
    orc.orthonormalize_left(cores, k)
for k in range(len(cores) - 1, 0, -1):
    tn.node_orthonormalize_right(tn.main_nodes[k])
    orc.orthonormalize_right(cores, k)
shapes = [tuple(n.tensor.shape) for n in tn.main_nodes]
for n, c in zip(tn.main_nodes, cores):
    g = n.tensor.cpu().numpy()
    assert g.shape == c.shape, (g.shape, c.shape)
    worst = max(worst, float(np.linalg.norm(g - c) / np.linalg.norm(c)))
p1 = tn.forward(x, to_tensor=True).cpu().numpy()
drift = float(np.linalg.norm(p1 - p0) / np.linalg.norm(p0))
print("qr_wide_check", shapes, "worst core err %.2e" % worst, "prediction drift %.2e" % drift, "ok" if worst < 1e-11 and drift < 1e-11 else "FAIL",
      "%.1f s" % (time.time() - t0), flush=True)
