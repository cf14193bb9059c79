"""Cum-sum train on the CPU stand-in kernels vs the recording of the reference's operator-node CumSumLayer."""
import torch

import cumsum_case
import fake_ops

torch.set_default_dtype(torch.float64)


def test_cumsum_train_cpu(monkeypatch):
    fake_ops.install(monkeypatch)
    cumsum_case.run("cpu")


def test_cum_sum_operator_tensor():
    """``get_cum_sum_operator`` (reference tensor/layers.py:408-423): contracting the operator train with per-site inputs gives the
    running sums x_1[k], x_1[k] + ... the closed form of tensor/cumsum.py is built on -- checked from the definition, no reference needed:
    (i, k, k, m) is one for i <= k and m = k; a single row on the first site; a single column on the last."""
    import torch
    from tensornetworksfork_b200.tensor.layers import get_cum_sum_operator
    f, N = 4, 3
    first, mid, last = (get_cum_sum_operator(n, N, f, dtype=torch.float64) for n in range(N))
    assert tuple(first.shape) == (1, f, f, f) and tuple(mid.shape) == (f, f, f, f) and tuple(last.shape) == (f, f, f, 1)
    for k in range(f):
        assert float(first[0, k, k, k]) == 1.0 and float(first[0, k, k].sum()) == 1.0
        for i in range(f):
            assert float(mid[i, k, k, k]) == (1.0 if i <= k else 0.0)
            assert float(last[i, k, k, 0]) == (1.0 if i <= k else 0.0)
    assert float(first.sum()) == f and float(mid.sum()) == f * (f + 1) / 2 and float(last.sum()) == f * (f + 1) / 2
    one = get_cum_sum_operator(0, 1, f, dtype=torch.float64)
    assert tuple(one.shape) == (1, f, f, 1) and float(one.sum()) == f
