"""Golden recording of the reference's CumSumLayer sweep -- build container only."""
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from tensor.layers import CumSumLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(21)
    N, F, r, n = 260, 3, 3, 3
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(X[:, :1] * X[:, 1:2] + 0.5 * X[:, 2:3] ** 2 + 0.3 * X[:, :1] + 0.05 * rng.normal(size=(N, 1)))
    torch.manual_seed(17)
    layer = CumSumLayer(n, r, F + 1, output_shape=1, constrict_bond=False, perturb=False)
    tn = layer.tensor_network
    flat = {"x": Xb.numpy(), "y": y.numpy(), "n_cores": np.array(n)}
    for i, nd in enumerate(tn.train_nodes):
        flat[f"cores0_{i}"] = nd.tensor.detach().numpy().copy()
    flat["pred0"] = tn.forward(Xb, to_tensor=True).detach().numpy().copy()
    ups = []
    orig_solve = tn.solve_system
    cur = {}

    def solve_system(node, A, b, method="exact", eps=0.0):
        cur["A"] = A.detach().numpy().copy()
        cur["b"] = b.detach().numpy().copy()
        cur["before"] = [t.tensor.detach().numpy().copy() for t in tn.train_nodes]
        return orig_solve(node, A, b, method=method, eps=eps)

    tn.solve_system = solve_system

    def block_callback(NS, node):
        u = dict(cur)
        cur.clear()
        u["NS"], u["k"] = NS, tn.train_nodes.index(node)
        u["after"] = [t.tensor.detach().numpy().copy() for t in tn.train_nodes]
        ups.append(u)

    tn.accumulating_swipe(Xb, y, SquareBregFunction(), batch_size=64, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5, eps_decay=0.5,
                          block_callback=block_callback, loss_callback=lambda NS, nd, l: cur.__setitem__("loss", float(l)))
    flat["n_updates"] = np.array(len(ups))
    for ui, u in enumerate(ups):
        flat[f"u{ui}_scal"] = np.array([u["NS"], u["k"], u["loss"]])
        flat[f"u{ui}_A"] = u["A"]
        flat[f"u{ui}_b"] = u["b"]
        for i in range(n):
            flat[f"u{ui}_before_{i}"] = u["before"][i]
            flat[f"u{ui}_after_{i}"] = u["after"][i]
    flat["pred"] = tn.forward_batch(Xb, 64).detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "cumsum_reg.npz"), **flat)
    print("cumsum_reg:", len(ups), "updates", [round(u["loss"], 6) for u in ups])


if __name__ == "__main__":
    main()
