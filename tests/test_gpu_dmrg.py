"""Growing 2-site flow on the real kernels (needs a B200).  Cores are compared through predictions only:
the host SVD of split_node fixes signs differently on CPU and GPU."""
import pytest
import torch

import dmrg_case

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def test_growing_dmrg_flow_gpu():
    dmrg_case.run("cuda", compare_cores=False)
