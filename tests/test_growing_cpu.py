"""Growing-TT estimator (reference tensor/module.py:502-614) host logic on the CPU stand-in kernels vs the reference recording."""
import pytest
import torch

import fake_ops
import growing_case as gc

torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_growing_tt_one_pass_schedule(tag, monkeypatch):
    fake_ops.install(monkeypatch)
    hist_err, pred_err, core_err, score_err = gc.run(tag, "cpu")
    assert hist_err < 1e-8 and pred_err < 1e-7 and core_err < 1e-6 and score_err < 1e-7, (hist_err, pred_err, core_err, score_err)


@pytest.mark.parametrize("tag", ["unique", "same", "block"])
def test_minibatch_estimator_swipe_methods(tag, monkeypatch):
    """TensorTrainBatchRegressor (reference tensor/module.py:308-500): shuffled epochs, one accumulating_swipe per minibatch, the
    three ways of assigning cores to minibatches; validation trajectory, predictions and cores against the recording."""
    import batch_case as bc
    fake_ops.install(monkeypatch)
    traj_err, pred_err, core_err = bc.run(tag, "cpu")
    assert traj_err < 1e-9 and pred_err < 1e-8 and core_err < 1e-7, (traj_err, pred_err, core_err)


def test_mirrored_cycle_order():
    from itertools import islice
    from tensornetworksfork_b200.tensor.module import mirrored_cycle
    assert list(mirrored_cycle([1, 2, 3, 4], one_cycle=True)) == [1, 2, 3, 4, 3, 2, 1]
    assert list(islice(mirrored_cycle([1, 2, 3], one_cycle=False), 9)) == [1, 2, 3, 2, 1, 2, 3, 2, 1]
    assert list(mirrored_cycle([], one_cycle=False)) == [] and list(mirrored_cycle([7], one_cycle=True)) == [7]
