"""Golden recording of the reference's sklearn-style wrappers (models/tensor_train.py:91-315 TensorTrainRegressor, models/tnml.py
TNMLRegressor): fit with validation split and EarlyStopping, predict, score -- build container only.

    python tests/golden/make_golden_wrappers.py        ->  tests/golden/wrappers.npz
"""
import json
import os
import sys
import types

import numpy as np

m = types.ModuleType("matplotlib"); p = types.ModuleType("matplotlib.pyplot"); m.pyplot = p
sys.modules["matplotlib"] = m; sys.modules["matplotlib.pyplot"] = p
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from models.tensor_train import TensorTrainRegressor  # noqa: E402  (the REFERENCE's classes)
from models.tnml import TNMLRegressor  # noqa: E402
from tensor.bregman import XEAutogradBregman  # noqa: E402
import wrappers_case as wc  # noqa: E402


def main():
    flat = {}
    for name in wc.ALL:
        cls = TensorTrainRegressor if name in wc.TT_CASES else TNMLRegressor
        pred, score, est = wc.fit(name, cls, XEAutogradBregman, "cpu")
        es = getattr(est, "_early_stopper", None)
        flat[f"{name}_pred"] = np.asarray(pred)
        flat[f"{name}_score"] = np.array(float(score))
        flat[f"{name}_n_val"] = np.array(len(es.val_history) if es is not None else -1)
        flat[f"{name}_meta"] = np.array(json.dumps({"kw": wc.TT_CASES.get(name) or wc.TNML_CASES[name]}))
        print(name, "score", float(score), "validation evaluations", int(flat[f"{name}_n_val"]), "pred", np.asarray(pred).shape)
    np.savez_compressed(os.path.join(HERE, "wrappers.npz"), **flat)


if __name__ == "__main__":
    main()
