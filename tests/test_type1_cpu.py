"""Type-I model (sum of tensor trains with 1..N cores) on the CPU stand-in kernels vs the reference recording."""
import torch

import fake_ops
import type1_case

torch.set_default_dtype(torch.float64)


def test_type1_sum_of_networks_cpu(monkeypatch):
    fake_ops.install(monkeypatch)
    type1_case.run("cpu")
