"""lanczos_swipe / scipy_swipe host logic against the reference recordings, on the CPU stand-in kernels."""
import pytest
import torch

import fake_ops
import krylov_cases as kc

torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("name", ["krylov_lanczos_reg", "krylov_lanczos_xe"])
def test_lanczos_swipe_with_injected_start_vector(name, monkeypatch):
    fake_ops.install(monkeypatch)
    core_err, loss_err = kc.run_case(name, "cpu")
    assert core_err < 1e-8 and loss_err < 1e-9, (core_err, loss_err)


@pytest.mark.parametrize("name", ["krylov_scipy_cg", "krylov_scipy_minres"])
def test_scipy_swipe_float32_host_recurrences(name, monkeypatch):
    """Passing the SciPy solver object reproduces the reference's float32 host recurrences (network.py:918-926)."""
    fake_ops.install(monkeypatch)
    core_err, loss_err = kc.run_case(name, "cpu", scipy_object=True)
    assert core_err < 5e-5 and loss_err < 5e-5, (core_err, loss_err)


@pytest.mark.parametrize("name", ["krylov_scipy_cg", "krylov_scipy_minres"])
def test_device_krylov_solvers_track_the_float32_reference(name, monkeypatch):
    """'cg' / 'minres' strings select the float64 on-device solvers.  The local systems carry no ridge and are singular by
    gauge freedom, so cores are not comparable with the float32 reference; the per-node losses are (5e-3)."""
    fake_ops.install(monkeypatch)
    core_err, loss_err = kc.run_case(name, "cpu", scipy_object=False)
    assert loss_err < 5e-3, (core_err, loss_err)


def test_cumsum_lanczos_swipe_with_injected_start_vector(monkeypatch):
    """Matrix-free sweep of the cum-sum train (the reference runs lanczos_swipe on the operator-node graph of CumSumLayer)."""
    fake_ops.install(monkeypatch)
    core_err, loss_err = kc.run_case("krylov_cumsum_lanczos", "cpu")
    assert core_err < 1e-8 and loss_err < 1e-9, (core_err, loss_err)


def test_cumsum_scipy_swipe_float32_host_recurrences(monkeypatch):
    fake_ops.install(monkeypatch)
    core_err, loss_err = kc.run_case("krylov_cumsum_cg", "cpu", scipy_object=True)
    assert core_err < 5e-4 and loss_err < 5e-5, (core_err, loss_err)


@pytest.mark.parametrize("fused_map", [True, False])
def test_baseline_config4a_local_size_chain_against_reference_recording(fused_map, monkeypatch):
    """BASELINE config 4a (TNML classifier: sin-cos map, 9 logits on the first core, cross-entropy, rank 38 -> local systems of up to
    2888 parameters, scipy_swipe(cg)) on a 16-site chain and 512 rows against a recording of the unmodified reference: the 32 per-node
    losses and the prediction (the 784-site chain takes the reference the better part of an hour per pass)."""
    import cfg4a_case as c4
    fake_ops.install(monkeypatch)
    loss_err, pred_err = c4.run("cpu", fused_map=fused_map)
    assert loss_err.max() < 1e-9 and pred_err < 1e-9, (loss_err.max(), pred_err)
