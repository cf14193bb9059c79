"""Growing-TT estimator on the real kernels vs the reference recording (needs a B200)."""
import pytest
import torch

import growing_case as gc

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_growing_tt_one_pass_schedule_gpu(tag):
    hist_err, pred_err, core_err, score_err = gc.run(tag, "cuda")
    assert hist_err < 1e-8 and pred_err < 1e-7 and core_err < 1e-6 and score_err < 1e-7, (hist_err, pred_err, core_err, score_err)
