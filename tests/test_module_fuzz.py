"""Randomised ``fit / predict / score`` of the mirrored tensor/module.py estimators (TensorTrainRegressor, TensorTrainRegressorEarlyStopping,
TensorTrainBatchRegressor; reference tensor/module.py:103-600) against the UNMODIFIED reference classes: seeded random sites, rank,
epsilon range, minibatch size, method, model type, linear projection, early-stopping patience and the three minibatch schedules.
Build container only (needs /root/reference).

750 such configurations were run once (WF_LO / WF_HI select the seed range): 743 agree to 1e-5; one disagreement was a real gap
(an epsilon list shorter than the trained-node list must only fail when the sweep gets there, test_host_logic.py); 5 are constricted
trains the reference itself cannot contract; one is a 32-row-minibatch schedule at ridge 0.01 whose second epoch amplifies rounding
differences to order one (two runs of the reference alone differ as much)."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tensor")), reason="reference tree not mounted")
torch.set_default_dtype(torch.float64)

def _ref(module):
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.append(REF)
    return importlib.import_module(module)

def _data(seed, N, F):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(N, F))
    y = (np.tanh(X @ rng.normal(size=(F, 1))) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(N, 1)))
    k = int(0.75 * N)
    return X[:k], y[:k], X[k:], y[k:]

def draw(seed):
    rng = np.random.default_rng(6000 + seed)
    cls = str(rng.choice(["plain", "early", "batch"]))
    kw = dict(N=int(rng.integers(2, 6)), r=int(rng.integers(2, 5)), seed=int(rng.integers(0, 100)), constrict_bond=bool(rng.integers(0, 2)),
              perturb=bool(rng.integers(0, 2)), eps_start=float(rng.choice([1.0, 0.5, 1e-1])), eps_end=float(rng.choice([1.0, 1e-1, 1e-2])),
              batch_size=int(rng.choice([32, 64, 500])), method=str(rng.choice(["ridge_cholesky", "ridge_exact"])))
    if cls != "early":
        kw["num_swipes"] = int(rng.integers(1, 4))
        kw["model_type"] = str(rng.choice(["tt", "tt", "cpd"]))
    if rng.integers(0, 4) == 0 and kw.get("model_type", "tt") == "tt":
        kw["linear_dim"] = int(rng.integers(2, 4))
    if cls == "early":
        kw["early_stopping"] = int(rng.choice([2, 3, 10]))
    if cls == "batch":
        kw["swipe_method"] = str(rng.choice(["batch_unique", "batch_same", "batch_block"]))
    return cls, kw, dict(N=int(rng.integers(150, 320)), F=int(rng.integers(3, 6)))

@pytest.mark.parametrize("seed", range(int(os.environ.get("WF_LO", 0)), int(os.environ.get("WF_HI", 12))))
def test_module_fuzz(seed, monkeypatch):
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_mod = _ref("tensor.module")
    from tensornetworksfork_b200.tensor import module as my_mod
    cls, kw, d = draw(seed)
    name = {"plain": "TensorTrainRegressor", "early": "TensorTrainRegressorEarlyStopping", "batch": "TensorTrainBatchRegressor"}[cls]
    Xtr, ytr, Xte, yte = _data(seed, d["N"], d["F"])
    out = []
    for mod in (ref_mod, my_mod):
        try:
            torch.manual_seed(5)
            np.random.seed(5)
            est = getattr(mod, name)(device="cpu", **kw)
            if cls == "early":
                est.fit(Xtr, ytr, X_val=Xte, y_val=yte)
            else:
                est.fit(Xtr, ytr)
            out.append(("ok", est.predict(Xte), est.score(Xte, yte)))
        except Exception as e:
            out.append(("exc", type(e).__name__, str(e)[:300]))
    r, m = out
    if r[0] == "exc":
        if r[1] == "RuntimeError" and "einsum" in r[2]:       # a constricted train the reference itself cannot contract
            return
        assert m[0] == "exc" and m[1] == r[1], (name, kw, r, m)  # e.g. the IndexError of an epsilon list that runs out mid-pass
        return
    assert m[0] == "ok", (name, kw, m)
    err = np.linalg.norm(np.asarray(m[1]) - np.asarray(r[1])) / max(np.linalg.norm(np.asarray(r[1])), 1e-12)
    assert err < 1e-5, (name, kw, d, err)
    assert abs(m[2] - r[2]) < 1e-5, (name, kw, m[2], r[2])
