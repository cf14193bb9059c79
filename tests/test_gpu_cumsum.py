"""Cum-sum train on the real kernels (needs a B200)."""
import pytest
import torch

import cumsum_case

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def test_cumsum_train_gpu():
    cumsum_case.run("cuda")
