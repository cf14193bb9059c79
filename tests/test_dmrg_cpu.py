"""Growing 2-site flow (grow_middle / block sweep / split_node) on the CPU stand-in kernels vs the reference recording."""
import torch

import dmrg_case
import fake_ops

torch.set_default_dtype(torch.float64)


def test_growing_dmrg_flow_cpu(monkeypatch):
    fake_ops.install(monkeypatch)
    dmrg_case.run("cpu", compare_cores=True)
