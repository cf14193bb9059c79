"""BASELINE config 5 in its TNML reading (5b) at full chain length and rank on a row subsample, recorded from the unmodified
reference (build container only):

    python tests/golden/make_golden_cfg5b.py        # several minutes

Synthetic higgs-shaped data: 28 features -> 28 sites, polynomial basis of degree 5 (physical dimension 6, models/tnml.py:18-23),
rank 38: TensorTrainLayer(28, 38, 6, constrict_bond=True) with QR re-gauging as models/tnml.py:149,218-227 runs it (bonds
1, 6, 36, 38, ..., 38, 36, 6, 1; largest local system P = 38 * 6 * 38 = 8664), ridge_cholesky, eps 1.0 * 0.5^NS, batch_size 512,
one full sweep = 55 site updates on 2048 rows (the configuration itself has 1M-10M rows).  opt_einsum stand-in as in
make_golden_cfg2.py.  The tests regenerate the data.
"""
import os
import sys
import types

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "tools"))
import ref_vs_port  # noqa: E402

oe_dir = "/tmp/tn_opt_einsum_standin"
os.makedirs(os.path.join(oe_dir, "opt_einsum"), exist_ok=True)
open(os.path.join(oe_dir, "opt_einsum", "__init__.py"), "w").write(ref_vs_port.STANDIN)
sys.path.insert(0, oe_dir)
m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

N, F, R, DEG = 2048, 28, 38, 5


def data():
    rng = np.random.default_rng(2027)
    X = rng.uniform(-1, 1, size=(N, F))
    W = rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1))
    return X, y


def main():
    assert torch.backends.opt_einsum.is_available()
    X, y = data()
    xs = [torch.tensor(np.stack([X[:, j] ** d for d in range(DEG + 1)], 1)) for j in range(F)]        # models/tnml.py:18-23
    layer = TensorTrainLayer(F, R, DEG + 1, output_shape=1, constrict_bond=True, seed=42)
    tn = layer.tensor_network
    tn.orthonormalize_left()
    trace = []
    ok = tn.accumulating_swipe(xs, torch.tensor(y), SquareBregFunction(), batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5, orthonormalize=True,
                               method="ridge_cholesky", num_swipes=1, skip_second=False, direction="l2r",
                               loss_callback=lambda NS, nd, l: (trace.append((NS, tn.train_nodes.index(nd), float(l))), print(trace[-1], flush=True)))
    pred = tn.forward([t[:256] for t in xs], to_tensor=True).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "cfg5b_chain28.npz"), ok=np.array(bool(ok)), trace=np.array(trace), pred256=pred, x_head=X[:4],
                        y_head=y[:4])
    print("ok", ok, len(trace), "updates; loss", trace[0][2], "->", trace[-1][2])


if __name__ == "__main__":
    main()
