"""Recordings of the reference for three more conv-TT shapes -- build container only.

    python tests/golden/make_golden_conv_type1.py

* conv_type1: the "type-I" image model of AAMNST.py:157-168 -- a SumOfNetworks of TensorConvolutionTrainLayers with 1..N columns,
  members after the first built without the bias patch / bias pixel -- under the dense accumulating_swipe of AAMNST.py:196-203;
* conv_onecol: a single column (the first member of such a sum) on its own;
* conv_nocb: convolution_bond = -1 (one pixel vector per column, no pixel bond).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_conv as mg  # noqa: E402  (stubs matplotlib, puts the reference on sys.path)
import torch  # noqa: E402
from tensor.layers import TensorConvolutionTrainLayer, TensorNetworkLayer  # noqa: E402
from tensor.network import SumOfNetworks  # noqa: E402
from tensor.bregman import SquareBregFunction, XEAutogradBregman  # noqa: E402


def main():
    torch.manual_seed(51)
    X, y = mg.data(11, 120, 5, 4, K=3)
    nets = [TensorConvolutionTrainLayer(num_carriages=i, bond_dim=3, num_patches=5 if i == 1 else 4, patch_pixels=4 if i == 1 else 3,
                                        output_shape=2, convolution_bond=2).tensor_network for i in range(1, 4)]
    layer = TensorNetworkLayer(SumOfNetworks(nets, train_operators=True))
    mg.record_dense("conv_type1", layer, X, y, XEAutogradBregman(w=1.0), batch_size=50, num_swipes=1, lr=1.0, method="ridge_cholesky",
                    eps=1.0, eps_decay=0.5)
    torch.manual_seed(52)
    X, y = mg.data(12, 90, 5, 4, C=2)
    layer = TensorConvolutionTrainLayer(num_carriages=1, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=2)
    mg.record_dense("conv_onecol", layer, X, y, SquareBregFunction(), batch_size=40, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=0.5,
                    eps_decay=0.5)
    torch.manual_seed(53)
    X, y = mg.data(13, 100, 5, 4, K=3)
    layer = TensorConvolutionTrainLayer(num_carriages=3, bond_dim=3, num_patches=5, patch_pixels=4, output_shape=2, convolution_bond=-1)
    mg.record_dense("conv_nocb", layer, X, y, XEAutogradBregman(w=1.0), batch_size=-1, num_swipes=1, lr=1.0, method="ridge_exact", eps=0.8,
                    eps_decay=0.5)


if __name__ == "__main__":
    main()
