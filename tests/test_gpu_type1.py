"""Type-I model on the real kernels (needs a B200)."""
import pytest
import torch

import type1_case

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def test_type1_sum_of_networks_gpu():
    type1_case.run("cuda")
