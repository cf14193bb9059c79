"""sklearn-style wrapper: poly-mode tensor train / CPD regressor and classifier.

Caller contract of the reference's ``TensorTrainRegressor`` (models/tensor_train.py:91-315) and its
``EarlyStopping`` (:29-89): same constructor arguments, ``fit / predict / score``, bias column
appended last, validation split by ``RandomState(seed).shuffle``, validation metric after every site
update, best weights restored after the fit.  The sweep itself runs on the B200 engine.
"""
from time import time

import numpy as np
import torch
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.metrics import accuracy_score, r2_score, root_mean_squared_error

from ..tensor.bregman import SquareBregFunction
from ..tensor.layers import CPDLayer, CumSumLayer, TensorNetworkLayer, TensorTrainLayer, TensorTrainLinearLayer
from ..tensor.module import unexplained_variance  # noqa: F401  (reference models/tensor_train.py:20-24, models/tnml.py:30-34)
from ..tensor.network import SumOfNetworks


def root_mean_squared_error_torch(y_true, y_pred):
    return root_mean_squared_error(y_true.cpu().numpy(), y_pred.cpu().numpy())


def error_rate_torch(y_true, y_pred):
    pred = torch.argmax(y_pred, dim=1).cpu().numpy()
    true = y_true.cpu().numpy()
    if true.ndim > 1 and true.shape[1] > 1:
        true = np.argmax(true, axis=1)
    return 1.0 - accuracy_score(true, pred)


class EarlyStopping:
    """Validation after every site update; keeps the best weights (reference models/tensor_train.py:29-89)."""

    def __init__(self, X_val, y_val, model_predict, get_model_weights=None, loss_fn=None, abs_err=0.0, rel_err=0.0,
                 early_stopping=5, verbose=0):
        self.X_val, self.y_val = X_val, y_val
        self.model_predict = model_predict
        self.get_model_weights = get_model_weights
        self.loss_fn = loss_fn
        self.abs_err, self.rel_err = abs_err, rel_err
        self.early_stopping = early_stopping
        self.verbose = verbose
        self.early_stop_count = 0
        self.best_val_loss = np.inf
        self.val_history, self.time_history = {}, {}
        w = self.get_model_weights() if get_model_weights is not None else None
        self.best_state_dict = w
        self.start_time = time()
        self.epoch = 0

    def convergence_criterion(self):
        elapsed = time() - self.start_time
        self.epoch += 1
        val = self.loss_fn(self.y_val, self.model_predict(self.X_val))
        self.val_history[self.epoch] = val
        self.time_history[self.epoch] = elapsed
        gain = self.best_val_loss - val
        good = gain >= self.abs_err or gain >= self.rel_err * abs(self.best_val_loss)
        if gain > 0:
            self.best_val_loss = val
            if self.get_model_weights is not None:
                self.best_state_dict = self.get_model_weights()
            self.early_stop_count = 0 if good else self.early_stop_count + 1
        else:
            self.early_stop_count += 1
        if self.early_stop_count >= self.early_stopping:
            if self.verbose > 0:
                print(f"Converged with best loss: {self.best_val_loss:.4f}")
            return True
        return False


def split_validation(X, y, seed, validation_split):
    """RandomState(seed).shuffle split (reference models/tensor_train.py:233-242)."""
    n = X.shape[0]
    idx = np.arange(n)
    np.random.RandomState(seed).shuffle(idx)
    cut = int(n * (1 - validation_split))
    tr, va = idx[:cut], idx[cut:]
    return X[tr], y[tr], X[va], y[va]


class TensorTrainRegressor(BaseEstimator, RegressorMixin):
    def __init__(self, N=3, r=8, output_dim=1, linear_dim=None, constrict_bond=False, perturb=False, seed=42,
                 device="cuda", bf=None, lr=1.0, eps_start=1.0, eps_decay=0.5, abs_err=1e-4, rel_err=1e-3, batch_size=512,
                 method="ridge_cholesky", num_swipes=30, model_type="tt", task="regression", train_operator=False,
                 cum_sum=False, early_stopping=0, verbose=0, gram_mode="fp64"):
        self.N, self.r, self.output_dim = N, r, output_dim
        self.linear_dim = linear_dim if linear_dim is not None and linear_dim > 0 else None
        self.constrict_bond, self.perturb, self.seed, self.device = constrict_bond, perturb, seed, device
        self.bf = bf if bf is not None else SquareBregFunction()
        self.lr, self.eps, self.eps_decay = lr, eps_start, eps_decay
        self.abs_err, self.rel_err = abs_err, rel_err
        self.batch_size, self.method, self.num_swipes = batch_size, method, num_swipes
        self.model_type, self.task, self.train_operator, self.cum_sum = model_type, task, train_operator, cum_sum
        self.early_stopping, self.verbose, self.gram_mode = early_stopping, verbose, gram_mode
        self._model = None
        if self.perturb and self.output_dim > 1:
            raise ValueError("perturb not supported for output dim > 1")

    def _initialize_model(self):
        if self.input_dim is None:
            raise ValueError("input_dim must be set")
        mt = self.model_type
        if "type1" in mt or "typeI" in mt:
            # sum of models with 1..N cores; members after the first do not see the bias column (reference :138-176)
            def member(i):
                f = self.input_dim - 1 if i != 1 else self.input_dim
                if mt.startswith("cpd"):
                    return CPDLayer(i, self.r, f, output_shape=self.output_dim, perturb=self.perturb, seed=self.seed + i)
                cls = CumSumLayer if self.cum_sum else TensorTrainLayer
                return cls(i, bond_dim=self.r, input_features=f, output_shape=self.output_dim, constrict_bond=self.constrict_bond,
                           perturb=self.perturb, seed=self.seed + i)
            if self.linear_dim is not None and self.linear_dim < self.input_dim:
                if not mt.startswith("tt") or self.cum_sum:
                    raise NotImplementedError("linear projections exist for tensor-train members only")
                nets = [TensorTrainLinearLayer(i, bond_dim=self.r, input_features=self.input_dim - 1 if i != 1 else self.input_dim,
                                               linear_dim=self.linear_dim, output_shape=self.output_dim,
                                               constrict_bond=self.constrict_bond, perturb=self.perturb,
                                               seed=self.seed + i).tensor_network for i in range(1, self.N + 1)]
            else:
                nets = [member(i).tensor_network for i in range(1, self.N + 1)]
            self._model = TensorNetworkLayer(SumOfNetworks(nets, output_labels=nets[0].output_labels,
                                                           train_operators=self.train_operator)).to(self.device)
            self._model.tensor_network.gram_mode = self.gram_mode
            return
        if self.linear_dim is not None and self.linear_dim < self.input_dim and mt.startswith("tt") and not self.cum_sum:
            # trainable linear projection in front of every core (reference models/tensor_train.py:191-197)
            self._model = TensorTrainLinearLayer(self.N, self.r, self.input_dim, self.linear_dim, output_shape=self.output_dim,
                                                 constrict_bond=self.constrict_bond, perturb=self.perturb, seed=self.seed).to(self.device)
            self._model.tensor_network.gram_mode = self.gram_mode
            return
        if mt.startswith("cpd"):
            self._model = CPDLayer(self.N, self.r, self.input_dim, output_shape=self.output_dim, perturb=self.perturb,
                                   seed=self.seed).to(self.device)
        elif mt.startswith("tt") and self.cum_sum:
            self._model = CumSumLayer(self.N, self.r, self.input_dim, output_shape=self.output_dim,
                                      constrict_bond=self.constrict_bond, perturb=self.perturb, seed=self.seed).to(self.device)
        elif mt.startswith("tt"):
            self._model = TensorTrainLayer(self.N, self.r, self.input_dim, output_shape=self.output_dim,
                                           constrict_bond=self.constrict_bond, perturb=self.perturb, seed=self.seed).to(self.device)
        else:
            raise ValueError(f"unknown model_type {mt!r}")
        self._model.tensor_network.gram_mode = self.gram_mode

    def _t(self, a):
        return torch.tensor(a, dtype=torch.float64, device=self.device) if isinstance(a, np.ndarray) else a

    def _with_bias(self, X):
        return torch.cat((X, torch.ones((X.shape[0], 1), dtype=torch.float64, device=X.device)), dim=1)

    def _predict_t(self, Xb):
        y = self._model.tensor_network.forward_batch(Xb, self.batch_size)
        if self.task == "classification":
            y = torch.cat([y, torch.zeros_like(y[..., :1])], dim=-1)
        return y

    def fit(self, X, y, X_val=None, y_val=None, validation_split=0.1, split_train=True):
        X, y = self._t(X), self._t(y)
        X = self._with_bias(X)
        if self._model is None:
            self.input_dim = X.shape[1]
            self._initialize_model()
        if X_val is None or y_val is None:
            if split_train:
                X_train, y_train, X_val, y_val = split_validation(X, y, self.seed, validation_split)
            else:
                X_train, y_train, X_val, y_val = X, y, X, y
        else:
            X_val, y_val = self._t(X_val), self._t(y_val)
            X_train, y_train = X, y
            if X_val.shape[1] != X_train.shape[1]:
                X_val = self._with_bias(X_val)
        self._early_stopper = EarlyStopping(
            X_val, y_val, model_predict=self._predict_t, get_model_weights=self._model.node_states,
            loss_fn=root_mean_squared_error_torch if self.task == "regression" else error_rate_torch,
            abs_err=self.abs_err, rel_err=self.rel_err, early_stopping=self.early_stopping, verbose=self.verbose)
        self._model.tensor_network.accumulating_swipe(
            X_train, y_train, self.bf, batch_size=self.batch_size, lr=self.lr, eps=self.eps, eps_decay=self.eps_decay,
            convergence_criterion=self._early_stopper.convergence_criterion, orthonormalize=False, method=self.method,
            verbose=self.verbose, num_swipes=self.num_swipes, skip_second=False, direction="l2r",
            disable_tqdm=self.verbose < 3)
        if self._early_stopper.best_state_dict is not None:
            self._model.load_node_states(self._early_stopper.best_state_dict, set_value=True)
        return self

    def predict(self, X):
        return self._predict_t(self._with_bias(self._t(X))).detach().cpu().numpy()

    def score(self, X, y_true):
        if not isinstance(y_true, np.ndarray):
            y_true = y_true.cpu().numpy()
        y_pred = self.predict(X)
        return r2_score(y_true, y_pred) if self.task == "regression" else accuracy_score(y_true, np.argmax(y_pred, axis=1))
