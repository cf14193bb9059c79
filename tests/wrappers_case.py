"""Shared body of the caller-level parity tests: ``fit / predict / score`` of the sklearn-style wrappers a user of the reference calls
(models/tensor_train.py:212-296, models/tnml.py:157-234) -- constructor, bias column, validation split, EarlyStopping restoring the
best weights -- against recordings of the UNMODIFIED reference classes on the same data and seed (tests/golden/wrappers.npz,
written by tests/golden/make_golden_wrappers.py)."""
import json
import os

import numpy as np
import torch

import golden_util as gu

TT_CASES = {
    "tt": dict(N=3, r=3, model_type="tt", num_swipes=3, eps_start=1.0, eps_decay=0.5, batch_size=64),
    "tt_perturb_earlystop": dict(N=4, r=3, perturb=True, model_type="tt", num_swipes=4, eps_start=0.5, eps_decay=0.7, batch_size=-1,
                                 early_stopping=3),
    "cpd": dict(N=3, r=4, model_type="cpd", num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=80),
    "tt_type1": dict(N=3, r=2, model_type="tt_type1", num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64, perturb=True),
    "cpd_type1": dict(N=3, r=3, model_type="cpd_type1", num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64),
    "tt_cumsum": dict(N=3, r=3, model_type="tt", cum_sum=True, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64),
    "tt_linear": dict(N=3, r=3, model_type="tt", linear_dim=2, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64),
    "tt_classifier": dict(N=3, r=3, model_type="tt", task="classification", output_dim=2, num_swipes=2, eps_start=1.0, eps_decay=0.5,
                          batch_size=64, xe=True),
    # early_stopping = 0 (the default) stops a fit after its first site update (reference models/tensor_train.py:84); the cases below
    # run whole sweeps under the validation callback and restore the best weights afterwards
    "tt_sweeps": dict(N=3, r=3, model_type="tt", num_swipes=3, eps_start=1.0, eps_decay=0.5, batch_size=64, early_stopping=50),
    "cpd_sweeps": dict(N=3, r=4, model_type="cpd", num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=80, early_stopping=50),
    "tt_type1_sweeps": dict(N=3, r=2, model_type="tt_type1", num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64, early_stopping=4),
    "tt_cumsum_sweeps": dict(N=3, r=3, model_type="tt", cum_sum=True, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64,
                             early_stopping=50),
    "tt_classifier_sweeps": dict(N=3, r=3, model_type="tt", task="classification", output_dim=2, num_swipes=2, eps_start=1.0,
                                 eps_decay=0.5, batch_size=64, xe=True, early_stopping=50),
}
TNML_CASES = {
    "tnml_sincos": dict(basis="sin-cos", r=4, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64),
    "tnml_polynomial_earlystop": dict(basis="polynomial", degree=2, r=3, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=-1,
                                      early_stopping=2),
    "tnml_sincos_classifier": dict(basis="sin-cos", r=3, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64, task="classification",
                                   output_dim=2, xe=True),
    "tnml_sincos_sweeps": dict(basis="sin-cos", r=4, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64, early_stopping=50),
    "tnml_classifier_sweeps": dict(basis="sin-cos", r=3, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64, task="classification",
                                   output_dim=2, xe=True, early_stopping=50),
}
ALL = sorted(TT_CASES) + sorted(TNML_CASES)


def data(name):
    xe = (TT_CASES.get(name) or TNML_CASES[name]).get("xe", False)
    rng = np.random.default_rng(5 if name in TT_CASES else 6)
    N, F = 260, (4 if name in TT_CASES else 5)
    X = rng.uniform(-1, 1, size=(N, F))
    if not xe:
        y = (0.5 * X[:, 0] - X[:, 1] * X[:, 2] + 0.3 * X[:, 3] ** 2 + 0.05 * rng.normal(size=N))[:, None]
    else:
        y = np.eye(3)[np.argmax(X @ rng.normal(size=(F, 3)), axis=1)]
    return X[:200], y[:200], X[200:], y[200:]


def fit(name, estimator_cls, xe_loss, device, **extra):
    """Fit the estimator of one case; returns (prediction on the held-out rows, score, estimator)."""
    kw = dict(TT_CASES.get(name) or TNML_CASES[name])
    xe = kw.pop("xe", False)
    Xtr, ytr, Xte, yte = data(name)
    est = estimator_cls(device=device, seed=7, bf=xe_loss(w=1.0) if xe else None, **kw, **extra)
    torch.manual_seed(99)          # CumSumLayer accepts `seed` without applying it (reference layers.py:425-433): same RNG state
    est.fit(Xtr, ytr, validation_split=0.2)
    yscore = np.argmax(yte, axis=1) if xe else yte
    return est.predict(Xte), est.score(Xte, yscore), est


def run(name, device, **extra):
    """(relative prediction error, absolute score error, number of validation evaluations matches) against the recording."""
    import tensornetworksfork_b200 as tnb
    from tensornetworksfork_b200.models import TensorTrainRegressor, TNMLRegressor
    z = np.load(os.path.join(gu.GOLDEN_DIR, "wrappers.npz"))
    meta = json.loads(str(z[f"{name}_meta"]))
    assert meta["kw"] == {k: v for k, v in (TT_CASES.get(name) or TNML_CASES[name]).items()}, "recording made with other arguments"
    cls = TensorTrainRegressor if name in TT_CASES else TNMLRegressor
    pred, score, est = fit(name, cls, tnb.XEAutogradBregman, device, **extra)
    ref_pred = z[f"{name}_pred"]
    assert pred.shape == ref_pred.shape
    pred_err = float(np.linalg.norm(pred - ref_pred) / np.linalg.norm(ref_pred))
    score_err = abs(float(score) - float(z[f"{name}_score"]))
    es = getattr(est, "_early_stopper", None)
    n_val = len(es.val_history) if es is not None else -1
    return pred_err, score_err, (n_val == int(z[f"{name}_n_val"]))
