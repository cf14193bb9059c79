"""sklearn-style wrapper: one tensor-train site per feature with a local feature map (TNML).

Caller contract of the reference's ``TNMLRegressor`` (models/tnml.py:103-261): ``fbasis`` /
``polynomial_basis`` maps, left-orthonormalised start, ``accumulating_swipe(..., orthonormalize=True)``.
With ``fused_map=True`` (default) the map is evaluated inside the kernels from the raw ``X`` and the
n per-site (N, f) tensors of the reference (:194-199) are never materialised; ``fused_map=False`` builds
them exactly as the reference does.
"""
import math

import numpy as np
import torch
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.metrics import accuracy_score, r2_score

from ..tensor.bregman import SquareBregFunction
from ..tensor.layers import TensorTrainLayer
from ..tensor.network import MappedInput
from .tensor_train import EarlyStopping, error_rate_torch, root_mean_squared_error_torch, split_validation, unexplained_variance  # noqa: F401


def fbasis(X):
    """[cos(pi x/2), sin(pi x/2)] per feature (reference models/tnml.py:11-16)."""
    return [torch.stack([torch.cos((0.5 * math.pi) * X[:, i]), torch.sin((0.5 * math.pi) * X[:, i])], dim=-1)
            for i in range(X.shape[-1])]


def polynomial_basis(X, degree=3):
    """[x^0 .. x^degree] per feature (reference models/tnml.py:18-23)."""
    return [torch.stack([X[:, i] ** d for d in range(degree + 1)], dim=-1) for i in range(X.shape[-1])]


class TNMLRegressor(BaseEstimator, RegressorMixin):
    def __init__(self, r=8, output_dim=1, seed=42, device="cuda", bf=None, lr=1.0, eps_start=1.0, eps_decay=0.5,
                 abs_err=1e-6, rel_err=1e-4, batch_size=512, method="ridge_cholesky", num_swipes=30, model_type="tt",
                 task="regression", train_operator=False, early_stopping=0, basis="sin-cos", degree=3, verbose=0,
                 constrict_bond=True, fused_map=True, gram_mode="fp64"):
        self.r, self.output_dim, self.seed, self.device = r, output_dim, seed, device
        self.input_dim = degree + 1 if basis == "polynomial" else 2
        self.constrict_bond, self.perturb = constrict_bond, False
        self.bf = bf if bf is not None else SquareBregFunction()
        self.lr, self.eps, self.eps_decay = lr, eps_start, eps_decay
        self.abs_err, self.rel_err = abs_err, rel_err
        self.batch_size, self.method, self.num_swipes = batch_size, method, num_swipes
        self.model_type, self.task, self.train_operator = model_type, task, train_operator
        self.early_stopping, self.basis, self.degree, self.verbose = early_stopping, basis, degree, verbose
        self.fused_map, self.gram_mode = fused_map, gram_mode
        self._model = None

    def _initialize_model(self):
        self._model = TensorTrainLayer(self.N, self.r, self.input_dim, output_shape=self.output_dim,
                                       constrict_bond=self.constrict_bond, perturb=self.perturb, seed=self.seed).to(self.device)
        self._model.tensor_network.gram_mode = self.gram_mode

    def _t(self, a):
        return torch.tensor(a, dtype=torch.float64, device=self.device) if isinstance(a, np.ndarray) else a

    def _map(self, X):
        if self.fused_map:
            return MappedInput(X.contiguous(), kind=self.basis, degree=self.degree)
        return fbasis(X) if self.basis == "sin-cos" else polynomial_basis(X, degree=self.degree)

    def _predict_t(self, Xm):
        y = self._model.tensor_network.forward_batch(Xm, self.batch_size)
        if self.task == "classification":
            y = torch.cat([y, torch.zeros_like(y[..., :1])], dim=-1)
        return y

    def fit(self, X, y, X_val=None, y_val=None, validation_split=0.1, split_train=True):
        X, y = self._t(X), self._t(y)
        if self._model is None:
            self.N = X.shape[1]
            self._initialize_model()
        if X_val is None or y_val is None:
            if split_train:
                X_train, y_train, X_val, y_val = split_validation(X, y, self.seed, validation_split)
            else:
                X_train, y_train, X_val, y_val = X, y, X, y
        else:
            X_val, y_val = self._t(X_val), self._t(y_val)
            X_train, y_train = X, y
        X_train, X_val = self._map(X_train), self._map(X_val)
        self._early_stopper = EarlyStopping(
            X_val, y_val, model_predict=self._predict_t, get_model_weights=self._model.node_states,
            loss_fn=root_mean_squared_error_torch if self.task == "regression" else error_rate_torch,
            abs_err=self.abs_err, rel_err=self.rel_err, early_stopping=self.early_stopping, verbose=self.verbose)
        self._model.tensor_network.orthonormalize_left()
        self._model.tensor_network.accumulating_swipe(
            X_train, y_train, self.bf, batch_size=self.batch_size, lr=self.lr, eps=self.eps, eps_decay=self.eps_decay,
            convergence_criterion=self._early_stopper.convergence_criterion, orthonormalize=True, method=self.method,
            verbose=self.verbose, num_swipes=self.num_swipes, skip_second=False, direction="l2r",
            disable_tqdm=self.verbose < 3)
        if self._early_stopper.best_state_dict is not None:
            self._model.load_node_states(self._early_stopper.best_state_dict, set_value=True)
        return self

    def predict(self, X):
        return self._predict_t(self._map(self._t(X))).detach().cpu().numpy()

    def score(self, X, y_true):
        if not isinstance(y_true, np.ndarray):
            y_true = y_true.cpu().numpy()
        y_pred = self.predict(X)
        return r2_score(y_true, y_pred) if self.task == "regression" else accuracy_score(y_true, np.argmax(y_pred, axis=1))
