"""Cum-sum train on the CPU stand-in kernels vs the recording of the reference's operator-node CumSumLayer."""
import torch

import cumsum_case
import fake_ops

torch.set_default_dtype(torch.float64)


def test_cumsum_train_cpu(monkeypatch):
    fake_ops.install(monkeypatch)
    cumsum_case.run("cpu")
