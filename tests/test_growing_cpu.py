"""Growing-TT estimator (reference tensor/module.py:502-614) host logic on the CPU stand-in kernels vs the reference recording."""
import pytest
import torch

import fake_ops
import growing_case as gc

torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_growing_tt_one_pass_schedule(tag, monkeypatch):
    fake_ops.install(monkeypatch)
    hist_err, pred_err, core_err, score_err = gc.run(tag, "cpu")
    assert hist_err < 1e-8 and pred_err < 1e-7 and core_err < 1e-6 and score_err < 1e-7, (hist_err, pred_err, core_err, score_err)
