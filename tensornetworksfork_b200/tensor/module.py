"""The "growing tensor train" estimators of the reference's tensor/module.py, on the B200 engine.

``TensorTrainRegressorEarlyStopping`` (reference tensor/module.py:502-614) is the schedule behind the paper_plot_* timing
tables: a perturb-initialised train (every core but the first is the identity on the bias feature, so a train of N cores
starts as a degree-1 model) is swept ONCE, left to right, with one epsilon per core (``eps_per_node``); after each core the
validation loss is evaluated, the polynomial degree reached so far is recorded, and the best weights are kept.  The caller
contract (constructor arguments, ``fit / predict / score``, ``_best_degree``, ``_singular``) is the reference's; the sweep is
``TensorNetwork.accumulating_swipe`` of this package.
"""
from functools import partial
from time import time

import numpy as np
import torch
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.metrics import r2_score, root_mean_squared_error

from .bregman import SquareBregFunction
from .layers import CPDLayer, TensorNetworkLayer, TensorTrainLayer, TensorTrainLinearLayer
from .network import SumOfNetworks


def root_mean_squared_error_torch(y_true, y_pred):
    return root_mean_squared_error(y_true.cpu().numpy(), y_pred.cpu().numpy())


def unexplained_variance(y_true, y_pred):
    """Mean over samples of the residual sum of squares across outputs over the total one about the column means (reference
    tensor/module.py:16-20)."""
    total = ((y_true - y_true.mean(dim=0, keepdim=True)) ** 2).sum(dim=1, keepdim=True)
    resid = ((y_true - y_pred) ** 2).sum(dim=1, keepdim=True)
    return (resid / total).mean().item()


class EarlyStopping:
    """Validation after every core of the one-pass sweep, indexed by the degree reached (reference tensor/module.py:22-101)."""

    def __init__(self, X_train, y_train, X_val, y_val, model_predict, get_model_weights=None, loss_fn=None, abs_err=0.0,
                 rel_err=0.0, early_stopping=5, verbose=0, start_degree=1):
        self.X_train, self.y_train, self.X_val, self.y_val = X_train, y_train, X_val, y_val
        self.model_predict, self.get_model_weights, self.loss_fn = model_predict, get_model_weights, loss_fn
        self.abs_err, self.rel_err, self.early_stopping, self.verbose = abs_err, rel_err, early_stopping, verbose
        self.early_stop_count = 0
        self.cur_degree = self.best_degree = start_degree
        self.best_val_loss = self.best_train_loss = np.inf
        self.val_history, self.time_history = {}, {}
        self.best_state_dict = self.get_model_weights()
        self.start_time = time()

    def convergence_criterion(self):
        elapsed = time() - self.start_time
        val = self.loss_fn(self.y_val, self.model_predict(self.X_val))
        self.val_history[self.cur_degree] = val
        self.time_history[self.cur_degree] = elapsed
        train = None
        if self.verbose > 0:
            train = self.loss_fn(self.y_train, self.model_predict(self.X_train))
            print(f"Degree {self.cur_degree}: Train loss: {train:.4f}, Val loss: {val:.4f}")
        gain = self.best_val_loss - val
        good = gain >= self.abs_err or gain >= self.rel_err * abs(self.best_val_loss)
        if gain > 0:
            self.best_val_loss = val
            if train is not None:
                self.best_train_loss = train
            self.best_degree = self.cur_degree
            if self.get_model_weights is not None:
                self.best_state_dict = self.get_model_weights()
            self.early_stop_count = 0 if good else self.early_stop_count + 1
        else:
            self.early_stop_count += 1
        if self.early_stop_count >= self.early_stopping:
            if self.verbose > 0:
                print(f"Converged degree: {self.best_degree} with best loss: {self.best_val_loss:.4f}")
            return True
        self.cur_degree += 1
        return False

    def best_summary(self):
        return {"best_degree": self.best_degree, "best_val_loss": self.best_val_loss, "best_train_loss": self.best_train_loss,
                "best_state_dict": self.best_state_dict}


class TensorTrainRegressor(BaseEstimator, RegressorMixin):
    """tensor/module.py:103-288: geometric epsilon list (one per half-sweep, or one per core for a single pass)."""

    def __init__(self, N=2, r=2, output_dim=1, linear_dim=None, constrict_bond=True, perturb=True, seed=42, device="cuda", bf=None,
                 lr=1.0, eps_start=1e-12, eps_end=1e-12, batch_size=512, method="ridge_cholesky", num_swipes=5, model_type="tt",
                 verbose=0, gram_mode="fp64"):
        self.N, self.r, self.output_dim = N, r, output_dim
        self.linear_dim = linear_dim if linear_dim is not None and linear_dim > 0 else None
        self.constrict_bond, self.perturb, self.seed, self.device = constrict_bond, perturb, seed, device
        self.bf = bf if bf is not None else SquareBregFunction()
        self.lr = lr
        if num_swipes > 1:
            self.epss = (np.geomspace(eps_start, eps_end, 2 * num_swipes).tolist() if eps_end != eps_start
                         else [eps_end] * (2 * num_swipes))
        else:
            self.epss = np.geomspace(eps_start, eps_end, N).tolist()
        self.batch_size, self.method, self.num_swipes, self.model_type, self.verbose = batch_size, method, num_swipes, model_type, verbose
        self.gram_mode = gram_mode
        self._model = None
        self.trajectory = []
        if self.perturb and self.output_dim > 1:
            raise ValueError("perturb not supported for output dim > 1")

    def _initialize_model(self):
        if self.input_dim is None:
            raise ValueError("input_dim must be set")
        use_linear = self.linear_dim is not None and self.linear_dim < self.input_dim
        if use_linear and (self.model_type == "cpd" or self.model_type.startswith("tt_type1")):
            raise NotImplementedError("type-I sums of linear-projection trains are not part of the B200 path")
        if use_linear:
            self._model = TensorTrainLinearLayer(self.N, self.r, self.input_dim, self.linear_dim, output_shape=self.output_dim,
                                                 constrict_bond=self.constrict_bond, perturb=self.perturb, seed=self.seed)
        elif self.model_type == "cpd":
            self._model = CPDLayer(self.N, self.r, self.input_dim, output_shape=self.output_dim, perturb=self.perturb, seed=self.seed)
        elif self.model_type.startswith("tt_type1"):
            nets = [TensorTrainLayer(i, bond_dim=self.r,
                                     input_features=self.input_dim - 1 if "bias_first" in self.model_type and i != 1 else self.input_dim,
                                     output_shape=self.output_dim, constrict_bond=self.constrict_bond, perturb=self.perturb,
                                     seed=self.seed + i).tensor_network for i in range(1, self.N + 1)]
            self._model = TensorNetworkLayer(SumOfNetworks(nets, output_labels=nets[0].output_labels))
        else:
            self._model = TensorTrainLayer(self.N, self.r, self.input_dim, output_shape=self.output_dim,
                                           constrict_bond=self.constrict_bond, perturb=self.perturb, seed=self.seed)
        self._model = self._model.to(self.device)
        self._model.tensor_network.gram_mode = self.gram_mode

    def _t(self, a, col=False):
        if isinstance(a, np.ndarray):
            a = torch.tensor(a, dtype=torch.float64, device=self.device)
        return a.unsqueeze(1) if col and a.ndim == 1 else a

    @staticmethod
    def _with_bias(X):
        return torch.cat((X, torch.ones((X.shape[0], 1), dtype=torch.float64, device=X.device)), dim=1)

    def _split(self, X, y, X_val, y_val, validation_split, split_train):
        """Tensors with the bias column, the model (built on first use) and the train / validation split of the reference's
        ``fit`` (tensor/module.py:186-229)."""
        X, y = self._with_bias(self._t(X)), self._t(y, col=True)
        if self._model is None:
            self.input_dim = X.shape[1]
            self._initialize_model()
        if X_val is None or y_val is None:
            if split_train:
                idx = np.arange(X.shape[0])
                np.random.RandomState(self.seed).shuffle(idx)
                cut = int(X.shape[0] * (1 - validation_split))
                return X[idx[:cut]], y[idx[:cut]], X[idx[cut:]], y[idx[cut:]]
            return X, y, X, y
        X_val, y_val = self._t(X_val), self._t(y_val, col=True)
        if X_val.shape[1] != X.shape[1]:
            X_val = self._with_bias(X_val)
        return X, y, X_val, y_val

    def _validate(self, X_val, y_val, epoch):
        """One trajectory entry: validation RMSE (and accuracy for several outputs), reference tensor/module.py:234-249."""
        log = {"epoch": epoch}
        pred = self._model.tensor_network.forward_batch(X_val, self.batch_size)
        log["val_rmse"] = root_mean_squared_error_torch(pred, y_val)
        if y_val.shape[1] > 1:
            log["val_accuracy"] = (torch.argmax(pred, dim=1) == torch.argmax(y_val, dim=1)).float().mean().item()
        if self.verbose > 0:
            print(", ".join(f"{k}: {v:.4f}" if isinstance(v, float) else f"{k}: {v}" for k, v in log.items()))
        self.trajectory.append(log)

    def fit(self, X, y, X_val=None, y_val=None, validation_split=0.1, split_train=True):
        X_train, y_train, X_val, y_val = self._split(X, y, X_val, y_val, validation_split, split_train)
        self.trajectory = []

        def convergence_criterion():
            self._validate(X_val, y_val, len(self.trajectory) + 1)
            return False

        self._model.tensor_network.accumulating_swipe(
            X_train, y_train, self.bf, batch_size=self.batch_size, lr=self.lr, eps=self.epss,
            convergence_criterion=convergence_criterion, orthonormalize=False, method=self.method, verbose=self.verbose,
            num_swipes=self.num_swipes, skip_second=False, direction="l2r", disable_tqdm=self.verbose < 3,
            eps_per_node=(self.num_swipes == 1) and (len(self.epss) == self.N))
        return self

    def predict(self, X):
        return self._model.tensor_network.forward_batch(self._with_bias(self._t(X)), self.batch_size).detach().cpu().numpy()

    def score(self, X, y_true):
        if not isinstance(y_true, np.ndarray):
            y_true = y_true.cpu().numpy()
        return r2_score(y_true, self.predict(X).squeeze())


class TensorTrainRegressorEarlyStopping(TensorTrainRegressor):
    """One left-to-right pass over a perturb-initialised train, one epsilon per core, early stopping on the validation loss after
    every core (= every polynomial degree); the best weights are restored.  Reference tensor/module.py:502-614."""

    def __init__(self, *args, early_stopping=10, rel_err=1e-12, abs_err=1e-13, validation_split=0.1, split_train=False, **kwargs):
        if "num_swipes" in kwargs and kwargs["num_swipes"] != 1:
            print("Warning: num_swipes is not set to 1 for early stopping. This setting will be overridden.")
        if "perturb" in kwargs and not kwargs["perturb"]:
            print("Warning: perturb is not set to True for early stopping. This setting will be overridden.")
        kwargs["num_swipes"] = 1
        kwargs["perturb"] = True
        super().__init__(*args, **kwargs)
        self.early_stopping, self.rel_err, self.abs_err = early_stopping, rel_err, abs_err
        self.validation_split, self.split_train = validation_split, split_train

    def fit(self, X, y, X_val=None, y_val=None):
        X, y = self._t(X), self._t(y, col=True)
        if X_val is None or y_val is None:
            if self.split_train:
                idx = np.arange(X.shape[0])
                np.random.RandomState(self.seed).shuffle(idx)
                cut = int(X.shape[0] * (1 - self.validation_split))
                X_train, y_train, X_val, y_val = X[idx[:cut]], y[idx[:cut]], X[idx[cut:]], y[idx[cut:]]
            else:
                X_train, y_train, X_val, y_val = X, y, X, y
        else:
            X_val, y_val = self._t(X_val), self._t(y_val, col=True)
            X_train, y_train = X, y
        X_train, X_val = self._with_bias(X_train), self._with_bias(X_val)
        if self._model is None:
            self.input_dim = X_train.shape[1]
            self._initialize_model()
        self._early_stopping = EarlyStopping(
            X_train, y_train, X_val, y_val,
            model_predict=partial(self._model.tensor_network.forward_batch, batch_size=self.batch_size),
            get_model_weights=lambda: self._model.node_states(), loss_fn=root_mean_squared_error_torch, abs_err=self.abs_err,
            rel_err=self.rel_err, early_stopping=self.early_stopping, verbose=self.verbose)
        converged = self._model.tensor_network.accumulating_swipe(
            X_train, y_train, self.bf, batch_size=self.batch_size, convergence_criterion=self._early_stopping.convergence_criterion,
            eps=self.epss, method=self.method, skip_second=True, lr=self.lr, orthonormalize=False, verbose=self.verbose,
            num_swipes=1, direction="l2r", disable_tqdm=self.verbose < 3, eps_per_node=True)
        best = self._early_stopping.best_summary()
        self._best_degree = best["best_degree"]
        self._singular = not converged
        if best["best_state_dict"] is not None:
            self._model.load_node_states(best["best_state_dict"], set_value=True)
        return self


def mirrored_cycle(seq, one_cycle=False):
    """first .. last .. second, first .. : the back-and-forth visiting order of the cores (reference tensor/module.py:290-306).
    ``one_cycle`` stops after one forward pass and the way back without the last element."""
    seq = list(seq)
    if not seq:
        return
    if one_cycle:
        yield from seq + seq[-2::-1]
        return
    pattern = seq + seq[-2:0:-1]
    while True:
        yield from pattern


class TensorTrainBatchRegressor(TensorTrainRegressor):
    """Stochastic variant: every minibatch of a shuffled epoch drives its own ``accumulating_swipe`` (reference
    tensor/module.py:308-500; train_mnist_batch.py:54-73).  ``swipe_method``:

    * ``'batch_unique'`` -- one core per minibatch, cores visited back and forth across minibatches;
    * ``'batch_same'``   -- a full ``num_swipes`` sweep of all cores on every minibatch, validation after each;
    * ``'batch_block'``  -- one core at a time, updated on every minibatch of the epoch in turn.

    Each call binds a new batch, so the environments are rebuilt per call; the per-call work is one small Gram + solve."""

    def __init__(self, *args, batch_size=1024, swipe_method="batch_unique", **kwargs):
        super().__init__(*args, batch_size=batch_size, **kwargs)
        self.swipe_method = swipe_method

    def fit(self, X, y, X_val=None, y_val=None, validation_split=0.1, split_train=True):
        X_train, y_train, X_val, y_val = self._split(X, y, X_val, y_val, validation_split, split_train)
        if self.verbose > 0:
            print("Number of parameters:", self._model.num_parameters())
        tn = self._model.tensor_network
        n_train = X_train.shape[0]
        bs = self.batch_size
        n_batches = (n_train + bs - 1) // bs
        self.trajectory = []
        epoch = [0]
        counter = [0]

        def end_of_epoch():
            # validation only when the running minibatch count closes an epoch (reference :363-384)
            if counter[0] % n_batches == 0:
                epoch[0] += 1
                self._validate(X_val, y_val, epoch[0])
            return False

        common = dict(batch_size=-1, lr=self.lr, eps=self.epss, orthonormalize=False, method=self.method, verbose=self.verbose,
                      skip_second=False, direction="l2r", disable_tqdm=self.verbose < 3, eps_per_node=len(self.epss) == self.N)
        rng = np.random.RandomState(self.seed)
        for _ in range(self.num_swipes):
            order = rng.permutation(n_train)
            batches = [order[lo:lo + bs] for lo in range(0, n_train, bs)]
            if self.swipe_method == "batch_unique":
                cores = mirrored_cycle(tn.train_nodes, one_cycle=False)
                for rows in batches:
                    counter[0] += 1
                    tn.accumulating_swipe(X_train[rows], y_train[rows], self.bf, node_order=[next(cores)], num_swipes=1,
                                          convergence_criterion=end_of_epoch, **common)
            elif self.swipe_method == "batch_same":
                for rows in batches:
                    counter[0] += 1
                    tn.accumulating_swipe(X_train[rows], y_train[rows], self.bf, num_swipes=self.num_swipes, **common)
                    epoch[0] += 1
                    self._validate(X_val, y_val, epoch[0])
            elif self.swipe_method == "batch_block":
                for core in mirrored_cycle(tn.train_nodes, one_cycle=True):
                    for rows in batches:
                        counter[0] += 1
                        tn.accumulating_swipe(X_train[rows], y_train[rows], self.bf, node_order=[core], num_swipes=1,
                                              convergence_criterion=end_of_epoch, **common)
            # any other value: the reference's loop does nothing either
        return self
