"""TensorTrainLinearLayer on the real kernels against recordings of the reference (needs a B200)."""
import pytest
import torch

import linear_cases as lc

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("name", sorted(lc.CASES))
def test_linear_layer_sweeps_gpu(name):
    init_err, fwd_err, core_err, loss_err, pred_err = lc.run_case(name, "cuda")
    assert init_err == 0.0
    tight = lc.CASES[name]["kind"] == "dense"
    assert fwd_err < 1e-12 and core_err < (1e-7 if tight else 1e-6) and loss_err < (1e-9 if tight else 1e-5) and pred_err < (1e-7 if tight else 1e-3), \
        (fwd_err, core_err, loss_err, pred_err)
