"""lanczos_swipe / scipy_swipe on the real matvec kernels against the reference recordings (needs a B200)."""
import pytest
import torch

import krylov_cases as kc

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("name", ["krylov_lanczos_reg", "krylov_lanczos_xe"])
def test_lanczos_swipe_gpu(name):
    core_err, loss_err = kc.run_case(name, "cuda")
    assert core_err < 1e-8 and loss_err < 1e-9, (core_err, loss_err)


@pytest.mark.parametrize("name", ["krylov_scipy_cg", "krylov_scipy_minres"])
def test_scipy_swipe_gpu(name):
    core_err, loss_err = kc.run_case(name, "cuda", scipy_object=True)
    assert core_err < 5e-5 and loss_err < 5e-5, (core_err, loss_err)
    core_err, loss_err = kc.run_case(name, "cuda", scipy_object=False)
    assert loss_err < 5e-3, (core_err, loss_err)
