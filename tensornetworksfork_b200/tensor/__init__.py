from .node import TensorNode  # noqa: F401
from .network import TensorNetwork, SumOfNetworks, MappedInput, sweep_schedule  # noqa: F401
from .cpd import CPDNetwork  # noqa: F401
from .layers import (TensorNetworkLayer, TensorTrainLayer, CPDLayer, MainNodeLayer, InputNodeLayer,  # noqa: F401
                     TensorTrainDMRGInfiLayer, CumSumLayer, TensorConvolutionTrainLayer, TensorTrainLinearLayer)
from .conv import ConvTrainNetwork  # noqa: F401
from .cumsum import CumSumNetwork  # noqa: F401
from .bregman import (BregFunction, SquareBregFunction, AutogradLoss, XEAutogradBregman, KLDivBregman, SoftmaxSquaredLoss,  # noqa: F401
                      BinaryKLDivBregman, AutogradBregman, UncertaintyAutogradLoss)
