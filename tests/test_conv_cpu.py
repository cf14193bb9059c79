"""Conv-TT (TensorConvolutionTrainLayer): the numpy oracle against the reference recordings, the constructor against the
reference's initial cores, and the engine's host logic on the CPU stand-in kernels."""
import numpy as np
import pytest
import torch

import conv_cases as cc
import fake_ops
import golden_util as gu
from oracle import conv_oracle as co

torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("name", ["conv_lanczos_xe", "conv_lanczos_reg"])
def test_conv_oracle_reproduces_reference_recording(name):
    case, fx = cc.CASES[name], cc.load(name)
    C = fx["pred0"].shape[1]
    A, Cc = co.canon_cores(fx["cores0"], fx["names"], C)
    assert gu.relerr(co.forward(A, Cc, fx["x"]), fx["pred0"]) < 1e-12
    trace = []
    kw = case["kw"]
    cores, losses = co.lanczos_swipe(fx["cores0"], fx["names"], C, fx["x"], fx["y"], case["oloss"], kw["batch_size"], kw["num_swipes"],
                                     kw["lr"], kw["max_iter"], kw["tol"], [u["x0"] for u in fx["updates"]], trace=trace)
    assert [(t["NS"], t["k"]) for t in trace] == [(u["NS"], u["k"]) for u in fx["updates"]]
    err = max(gu.relerr(c, r) for t, u in zip(trace, fx["updates"]) for c, r in zip(t["after"], u["after"]))
    assert err < 1e-8, err
    assert np.max(np.abs(np.array(losses) - fx["losses"])) < 1e-10


def test_conv_constructor_matches_reference_shapes_and_order():
    for name in cc.CASES:
        cc.build(name, "cpu")          # asserts node names, order and shapes against the recording


@pytest.mark.parametrize("name", ["conv_lanczos_xe", "conv_lanczos_reg"])
@pytest.mark.parametrize("chunk", [None, 37])
def test_conv_lanczos_swipe_host_logic(name, chunk, monkeypatch):
    fake_ops.install(monkeypatch)
    fwd, core, loss, pred = cc.run_case(name, "cpu", chunk_rows=chunk)
    assert fwd < 1e-12 and core < 1e-8 and loss < 1e-9 and pred < 1e-8, (fwd, core, loss, pred)


@pytest.mark.parametrize("name", ["conv_scipy_cg", "conv_scipy_minres", "conv_scipy_cg_2col"])
def test_conv_scipy_swipe_host_logic(name, monkeypatch):
    fake_ops.install(monkeypatch)
    fwd, core, loss, pred = cc.run_case(name, "cpu", scipy_object=True)
    assert fwd < 1e-12 and core < 5e-4 and loss < 5e-5, (fwd, core, loss, pred)


@pytest.mark.parametrize("name", ["conv_scipy_cg", "conv_scipy_minres"])
def test_conv_device_krylov_solvers_track_the_float32_reference(name, monkeypatch):
    fake_ops.install(monkeypatch)
    fwd, core, loss, pred = cc.run_case(name, "cpu", scipy_object=False, loss_prefix=8)
    assert loss < 5e-3, (core, loss)


def test_conv_dense_sweep_is_refused(monkeypatch):
    fake_ops.install(monkeypatch)
    case, fx, layer = cc.build("conv_scipy_cg_2col", "cpu")
    with pytest.raises(NotImplementedError):
        layer.tensor_network.accumulating_swipe(torch.tensor(fx["x"]), torch.tensor(fx["y"]), case["loss"]())
