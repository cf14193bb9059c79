"""Recordings of the reference's matrix-free sweeps on the cum-sum train (CumSumLayer under lanczos_swipe / scipy_swipe,
tensor/network.py:709-932 on the operator-node graph of tensor/layers.py:425-477) -- build container only.

    python tests/golden/make_golden_cumsum_krylov.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_krylov as mk  # noqa: E402  (stubs matplotlib, puts the reference on sys.path)
import torch  # noqa: E402
from tensor.layers import CumSumLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402


def main():
    X, y = mk.data(5, 170, 3)
    torch.manual_seed(27)
    layer = CumSumLayer(3, 3, 4, output_shape=1, constrict_bond=False, perturb=False)
    mk.record("lanczos", "krylov_cumsum_lanczos", layer, X, y, SquareBregFunction(), batch_size=60, num_swipes=2, lr=1.0, max_iter=6, tol=1e-12)
    X, y = mk.data(6, 150, 3)
    torch.manual_seed(28)
    layer = CumSumLayer(4, 2, 4, output_shape=1, constrict_bond=False, perturb=False)
    mk.record("scipy", "krylov_cumsum_cg", layer, X, y, SquareBregFunction(), solver="cg", batch_size=-1, num_swipes=2, lr=1.0, max_iter=8,
              tol=1e-5)


if __name__ == "__main__":
    main()
