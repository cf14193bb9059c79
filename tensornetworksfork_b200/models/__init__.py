from .tensor_train import TensorTrainRegressor, EarlyStopping  # noqa: F401
from .tnml import TNMLRegressor, fbasis, polynomial_basis  # noqa: F401
