"""fit / predict / score of the mirrored sklearn-style wrappers (models/tensor_train.py:212-296, models/tnml.py:157-234 of the
reference) against the recording of the unmodified reference classes (tests/golden/wrappers.npz): the CPU twin runs on the stand-in
kernels, the GPU test drives the real kernels through the C ABI -- fit -> EarlyStopping -> load_node_states on the device."""
import pytest
import torch

import wrappers_case as wc

torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("name", wc.ALL)
def test_wrappers_follow_the_reference_recording_on_standin_kernels(name, monkeypatch):
    import fake_ops
    fake_ops.install(monkeypatch)
    pred_err, score_err, n_val_ok = wc.run(name, "cpu")
    assert pred_err < 1e-6 and score_err < 1e-6 and n_val_ok, (pred_err, score_err, n_val_ok)


@pytest.mark.gpu
@pytest.mark.parametrize("name", wc.ALL)
def test_wrappers_follow_the_reference_recording_gpu(name):
    pred_err, score_err, n_val_ok = wc.run(name, "cuda")
    assert pred_err < 1e-6 and score_err < 1e-6 and n_val_ok, (pred_err, score_err, n_val_ok)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tt_sweeps", "tt_perturb_earlystop", "cpd_sweeps", "tt_type1_sweeps", "tt_linear", "tt_classifier_sweeps"])
def test_wrappers_in_the_tensor_core_gram_mode_gpu(name):
    """The same fits with gram_mode='tf32x3': the Gram only preconditions the exact refinement, so the recording is followed at the
    same tolerance (cum-sum and the TNML wrappers have no gram_mode argument and are covered above)."""
    pred_err, score_err, n_val_ok = wc.run(name, "cuda", gram_mode="tf32x3")
    assert pred_err < 1e-6 and score_err < 1e-6 and n_val_ok, (pred_err, score_err, n_val_ok)
