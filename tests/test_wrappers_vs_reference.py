"""The sklearn-style wrappers a user of the reference calls (models/tensor_train.py, models/tnml.py): ``fit / predict / score`` of
the mirrored classes side by side with the UNMODIFIED reference classes, same arguments, same data, same seed -- the mirrored ones
on the CPU stand-in kernels.  Build container only (needs /root/reference); skipped on the GPU box."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not mounted")
torch.set_default_dtype(torch.float64)


def _ref(module):
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.append(REF)
    return importlib.import_module(module)


def _data(seed, N=260, F=4, classes=None):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(N, F))
    if classes is None:
        y = (0.5 * X[:, 0] - X[:, 1] * X[:, 2] + 0.3 * X[:, 3] ** 2 + 0.05 * rng.normal(size=N))[:, None]
    else:
        y = np.eye(classes)[np.argmax(X @ rng.normal(size=(F, classes)), axis=1)]
    return X[:200], y[:200], X[200:], y[200:]


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
}


@pytest.mark.parametrize("name", sorted(TT_CASES))
def test_tensor_train_regressor_fit_predict_score(name, monkeypatch):
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_mod = _ref("models.tensor_train")
    ref_breg = _ref("tensor.bregman")
    from tensornetworksfork_b200.models import TensorTrainRegressor
    import tensornetworksfork_b200 as tnb
    kw = dict(TT_CASES[name])
    xe = kw.pop("xe", False)
    Xtr, ytr, Xte, yte = _data(5, classes=3 if xe else None)
    out = []
    for cls, breg in ((ref_mod.TensorTrainRegressor, ref_breg), (TensorTrainRegressor, tnb)):
        bf = breg.XEAutogradBregman(w=1.0) if xe else None
        est = cls(device="cpu", seed=7, bf=bf, **kw)
        torch.manual_seed(99)          # CumSumLayer accepts `seed` without applying it (reference layers.py:425-433): same RNG state
        est.fit(Xtr, ytr, validation_split=0.2)
        yscore = np.argmax(yte, axis=1) if xe else yte
        out.append((est.predict(Xte), est.score(Xte, yscore), [n.tensor.detach().numpy() for n in est._model.tensor_network.train_nodes]))
    (rp, rs, rc), (mp, ms, mc) = out
    assert mp.shape == rp.shape
    assert np.linalg.norm(mp - rp) / np.linalg.norm(rp) < 1e-6, name
    assert abs(ms - rs) < 1e-6
    assert len(mc) == len(rc) and all(a.shape == b.shape for a, b in zip(mc, rc))


@pytest.mark.parametrize("kw", [dict(basis="sin-cos", r=4, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64),
                                dict(basis="polynomial", degree=2, r=3, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=-1, early_stopping=2),
                                dict(basis="sin-cos", r=3, num_swipes=2, eps_start=1.0, eps_decay=0.5, batch_size=64, task="classification",
                                     output_dim=2, xe=True)],
                         ids=["sincos", "polynomial_earlystop", "sincos_classifier"])
def test_tnml_regressor_fit_predict_score(kw, monkeypatch):
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_mod = _ref("models.tnml")
    ref_breg = _ref("tensor.bregman")
    from tensornetworksfork_b200.models import TNMLRegressor
    import tensornetworksfork_b200 as tnb
    kw = dict(kw)
    xe = kw.pop("xe", False)
    Xtr, ytr, Xte, yte = _data(6, F=5, classes=3 if xe else None)
    out = []
    for cls, breg in ((ref_mod.TNMLRegressor, ref_breg), (TNMLRegressor, tnb)):
        bf = breg.XEAutogradBregman(w=1.0) if xe else None
        est = cls(device="cpu", seed=7, bf=bf, **kw)
        est.fit(Xtr, ytr, validation_split=0.2)
        yscore = np.argmax(yte, axis=1) if xe else yte
        out.append((est.predict(Xte), est.score(Xte, yscore)))
    (rp, rs), (mp, ms) = out
    assert mp.shape == rp.shape
    assert np.linalg.norm(mp - rp) / np.linalg.norm(rp) < 1e-6
    assert abs(ms - rs) < 1e-6
