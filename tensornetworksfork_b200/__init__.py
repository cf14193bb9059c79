"""B200-native per-site Gauss-Newton / ALS sweep for tensor-train and CPD models.

Drop-in for the sweep path of niccogc/TensorNetworksFork: same layer constructors, the same
``TensorNetwork`` method set and the same sklearn-style wrappers, backed by hand-written sm_100a CUDA
kernels behind a C ABI (include/tn_b200.h).  See DESIGN.md and INTEGRATION.md.
"""
from . import ops  # noqa: F401
from .tensor import (TensorNode, TensorNetwork, SumOfNetworks, CPDNetwork, MappedInput, TensorNetworkLayer, TensorTrainLayer,  # noqa: F401
                     CPDLayer, TensorTrainDMRGInfiLayer, CumSumLayer, TensorConvolutionTrainLayer, TensorTrainLinearLayer, ConvTrainNetwork, BregFunction, SquareBregFunction, AutogradLoss, XEAutogradBregman, KLDivBregman, SoftmaxSquaredLoss,
                     BinaryKLDivBregman, AutogradBregman, UncertaintyAutogradLoss)

__version__ = "0.1.0"
