"""BASELINE config 3 at its full chain length on a row subsample, recorded from the unmodified reference (build container only):

    python tests/golden/make_golden_cfg3.py        # a few minutes

TNML regression with the sin-cos map on synthetic year-shaped data: 90 features -> 90 sites of physical dimension 2, rank 24,
TensorTrainLayer(90, 24, 2, constrict_bond=True) with QR re-gauging, exactly as models/tnml.py:149,218-227 runs it (bonds
1,2,4,8,16,24,...,24,16,8,4,2,1; ridge_cholesky, eps 1.0 * 0.5^NS, batch_size 512), one full sweep = 179 site updates.  The rows
are a 4096-row subsample of the 515k of the full configuration (the reference needs about an hour per sweep at full N even on its
authors' einsum path, used here through the opt_einsum stand-in of tools/ref_vs_port.py).  The tests regenerate the data.
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

N, F, R = 4096, 90, 24


def data():
    rng = np.random.default_rng(2026)
    X = rng.uniform(-1, 1, size=(N, F))
    W = rng.normal(size=(F, 1)) / np.sqrt(F)
    y = np.tanh(X @ W) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1))
    return X, y


def main():
    assert torch.backends.opt_einsum.is_available()
    X, y = data()
    xs = [torch.tensor(np.stack([np.cos(0.5 * np.pi * X[:, j]), np.sin(0.5 * np.pi * X[:, j])], 1)) for j in range(F)]   # models/tnml.py:11-16
    layer = TensorTrainLayer(F, R, 2, output_shape=1, constrict_bond=True, seed=42)
    tn = layer.tensor_network
    tn.orthonormalize_left()                                   # models/tnml.py:218
    trace = []
    ok = tn.accumulating_swipe(xs, torch.tensor(y), SquareBregFunction(), batch_size=512, lr=1.0, eps=1.0, eps_decay=0.5, orthonormalize=True,
                               method="ridge_cholesky", num_swipes=1, skip_second=False, direction="l2r",
                               loss_callback=lambda NS, nd, l: trace.append((NS, tn.train_nodes.index(nd), float(l))))
    pred = tn.forward([t[:256] for t in xs], to_tensor=True).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "cfg3_chain90.npz"), ok=np.array(bool(ok)), trace=np.array(trace), pred256=pred, x_head=X[:4],
                        y_head=y[:4])
    print("ok", ok, len(trace), "updates; loss", trace[0][2], "->", trace[-1][2])


if __name__ == "__main__":
    main()
