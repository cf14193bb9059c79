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
    assert loss < 2e-2, (core, loss)


@pytest.mark.parametrize("name", ["conv_type1", "conv_onecol", "conv_nocb"])
@pytest.mark.parametrize("chunk", [None, 37])
def test_conv_type1_and_degenerate_columns_host_logic(name, chunk, monkeypatch):
    """Type-I image model (sum of conv-TTs with 1..3 columns, AAMNST.py:157-203), a single column on its own, and
    convolution_bond = -1 (pixel vectors), each against its reference recording under the dense sweep."""
    fake_ops.install(monkeypatch)
    fwd, core, loss, pred = cc.run_case(name, "cpu", chunk_rows=chunk)
    assert fwd < 1e-12 and core < 1e-8 and loss < 1e-10 and pred < 1e-8, (fwd, core, loss, pred)


@pytest.mark.parametrize("name", ["conv_dense_xe", "conv_dense_reg"])
@pytest.mark.parametrize("chunk_bytes", [1 << 30, 20000])
def test_conv_dense_sweep_host_logic(name, chunk_bytes, monkeypatch):
    """accumulating_swipe on the conv layer (image_convolution_MNIST.py:120 call shape) against the reference recording; the small
    byte cap forces several Jacobian chunks per node."""
    fake_ops.install(monkeypatch)
    from tensornetworksfork_b200.tensor import conv
    monkeypatch.setattr(conv.ConvTrainNetwork, "dense_chunk_bytes", chunk_bytes, raising=False)
    fwd, core, loss, pred = cc.run_case(name, "cpu")
    assert fwd < 1e-12 and core < 1e-8 and loss < 1e-10 and pred < 1e-8, (fwd, core, loss, pred)


def test_grow_cart_reproduces_reference_cores_bit_for_bit():
    """Same seed, same constructor, two growths (reference tensor/layers.py:892-947): names, label order, shapes and values of all
    train nodes, including the random draw of the new pixel cores."""
    import tensornetworksfork_b200 as tnb
    z = cc.load_grow()
    torch.manual_seed(33)
    layer = tnb.TensorConvolutionTrainLayer(**cc.GROW_CTOR)
    layer.grow_cart(3, 2)
    layer.grow_cart()
    tn = layer.tensor_network
    assert [n.name for n in tn.train_nodes] == [str(s) for s in z["seed_names"]]
    assert [n.name for n in tn.main_nodes] == ["A1", "A2", "A3", "A4"] and len(tn.input_nodes) == 4
    assert layer.num_carriages == 4
    for i, n in enumerate(tn.train_nodes):
        ref = z[f"seed_core_{i}"]
        assert tuple(n.tensor.shape) == ref.shape, n.name
        assert np.array_equal(n.tensor.numpy(), ref), n.name


def test_conv_growing_flow_host_logic(monkeypatch):
    """Sweep, grow_cart, sweep 'r2l', grow_cart, minibatched sweep against the reference recording (call shape of
    image_convolution_growing_MNIST.py:84-103) on the CPU stand-in kernels."""
    fake_ops.install(monkeypatch)
    for pi, (fwd, core, loss, pred) in enumerate(cc.run_grow("cpu")):
        assert fwd < 1e-12 and core < 1e-8 and loss < 1e-10 and pred < 1e-8, (pi, fwd, core, loss, pred)


def _conv_shard_worker(rank, world, port, name, out_dir):
    import os
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops as fo
    import conv_cases as c2
    fo.install()
    case, fx, layer = c2.build(name, "cpu")
    tn = layer.tensor_network
    N = fx["y"].shape[0]
    cut = [0, N // 2 + 5, N][rank:rank + 2]           # uneven shards on purpose
    X, y = torch.tensor(fx["x"][cut[0]:cut[1]]), torch.tensor(fx["y"][cut[0]:cut[1]])
    tn.process_group, tn.shard_offset, tn.shard_total = dist.group.WORLD, cut[0], N
    x0s = [u["x0"] for u in fx["updates"]]
    cnt = [0]

    def x0_fn(node, b):
        v = torch.tensor(x0s[cnt[0]])
        cnt[0] += 1
        return v.reshape(-1)

    losses = []
    ok = tn.lanczos_swipe(X, y, case["loss"](), loss_callback=losses.append, x0_fn=x0_fn, **case["kw"])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ok=ok, losses=np.array(losses),
             **{f"core{i}": n.tensor.numpy() for i, n in enumerate(tn.train_nodes)})
    dist.destroy_process_group()


def test_conv_sample_sharded_lanczos_world2_gloo(tmp_path):
    """Two ranks with uneven row shards: the right-hand side and every matvec are sum-all-reduced; the result equals the
    reference recording and the cores are bit-identical on both ranks."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    name = "conv_lanczos_xe"
    mp.spawn(_conv_shard_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    fx = cc.load(name)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert bool(r0["ok"]) and bool(r1["ok"])
    for i in range(len(fx["cores0"])):
        assert np.array_equal(r0[f"core{i}"], r1[f"core{i}"]), "ranks diverged"
        assert gu.relerr(r0[f"core{i}"], fx["updates"][-1]["after"][i]) < 1e-7
    assert np.max(np.abs(r0["losses"] - fx["losses"])) < 1e-9


def test_baseline_config4b_full_model_size_against_reference_recording(monkeypatch):
    """BASELINE config 4 in its conv-TT reading at full model size (50 patches x 17 pixels, r = 38, CB = 4, 9 logits; the middle patch
    core has 72 200 parameters) under scipy_swipe(minres), 256 rows, against a recording of the unmodified reference: same initial
    cores from the same seed, the six per-node losses and the prediction."""
    import cfg4b_case as c4
    fake_ops.install(monkeypatch)
    loss_err, pred_err = c4.run("cpu")
    assert loss_err.max() < 1e-9 and pred_err < 1e-9, (loss_err, pred_err)
