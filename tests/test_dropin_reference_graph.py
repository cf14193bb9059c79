"""Drop-in check (build container only: needs /root/reference): the B200 engine class accepts the node graph built by
the REFERENCE's own layer constructors, and -- driven on the CPU stand-in kernels -- reproduces the reference's sweep.
Skipped where the reference tree is absent (the GPU box)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tensor")), reason="reference tree not mounted")

torch.set_default_dtype(torch.float64)


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)          # tensor/utils.py:2 imports pyplot; plotting unused
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.append(REF)
    import importlib
    ref_layers = importlib.import_module("tensor.layers")
    ref_breg = importlib.import_module("tensor.bregman")
    return ref_layers, ref_breg


def test_engine_runs_reference_built_graph(monkeypatch):
    import fake_ops
    ref_layers, ref_breg = _import_reference()
    from tensornetworksfork_b200.tensor.network import TensorNetwork as FastTN
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(0)
    N, F = 200, 4
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1].numpy()) + 0.1 * rng.normal(size=(N, 1)))
    kw = dict(batch_size=64, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)

    ref_layer = ref_layers.TensorTrainLayer(4, 3, F + 1, output_shape=1, constrict_bond=False, seed=5)
    ref_losses = []
    ref_layer.tensor_network.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: ref_losses.append(l), **kw)
    ref_pred = ref_layer.tensor_network.forward_batch(X, 64)

    layer2 = ref_layers.TensorTrainLayer(4, 3, F + 1, output_shape=1, constrict_bond=False, seed=5)
    old = layer2.tensor_network
    fast = FastTN(old.input_nodes, old.main_nodes, old.train_nodes, output_labels=old.output_labels, sample_dim=old.sample_dim)
    layer2.set_tensor_network(fast)                     # the reference's own layer object now owns the B200 engine
    losses = []
    ok = fast.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: losses.append(l), **kw)
    assert ok
    assert len(losses) == len(ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b))
    pred = fast.forward_batch(X, 64)
    assert float((pred - ref_pred).norm() / ref_pred.norm()) < 1e-8
    states = layer2.node_states()                       # reference's checkpoint hooks still work on the new engine
    layer2.load_node_states(states, set_value=True)
    assert float((fast.forward_batch(X, 64) - ref_pred).norm() / ref_pred.norm()) < 1e-8


def _pair(seed=7, n=4, r=3, F=3, C=1):
    """The same model twice: the reference's layer with its own engine, and the mirrored layer on the stand-in kernels."""
    ref_layers, ref_breg = _import_reference()
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(seed)
    N = 150
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :C].numpy()) + 0.1 * rng.normal(size=(N, C)))
    ref = ref_layers.TensorTrainLayer(n, r, F + 1, output_shape=C, constrict_bond=False, seed=seed)
    mine = tnb.TensorTrainLayer(n, r, F + 1, output_shape=C, constrict_bond=False, seed=seed)
    return ref, mine, X, y, ref_breg.SquareBregFunction(), tnb.SquareBregFunction()


def _trace(tn, X, y, loss, **kw):
    """(return value, [(callback kind, NS, node index)], final prediction) of one accumulating_swipe call."""
    ev = []
    idx = lambda node: tn.train_nodes.index(node)
    ret = tn.accumulating_swipe(X, y, loss, loss_callback=lambda NS, nd, l: ev.append(("loss", NS, idx(nd), round(float(l), 9))),
                                block_callback=lambda NS, nd: ev.append(("block", NS, idx(nd))), **kw)
    return ret, ev, tn.forward(X, to_tensor=True)


@pytest.mark.parametrize("kw", [
    dict(num_swipes=0),
    dict(num_swipes=2, skip_second=True, eps=[0.5, 0.25]),
    dict(num_swipes=1, direction="r2l", eps=[0.5, 0.25]),
    dict(num_swipes=1, node_order="middle_two"),
    dict(num_swipes=1, node_order="tuple"),
    dict(num_swipes=1, timeout=0.0),
    dict(num_swipes=2, stop_after=3),
    # becomes 'exact' (reference network.py:478-479): unregularised, nearly singular systems solved by LU in both engines, so only
    # the control flow and the first losses are comparable
    dict(num_swipes=1, method="ridge_exact", eps=0, loose=True),
    dict(num_swipes=1, batch_size=1000),                       # larger than the data set: one batch
    dict(num_swipes=1, batch_size=1, update_or_reset_stack="update"),
], ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items()))
def test_sweep_control_flow_matches_live_reference(kw, monkeypatch):
    """Return value, order and arguments of the callbacks, and the fitted function for the keyword combinations of SURVEY.md
    Appendix D that steer the sweep rather than the arithmetic (side by side with the reference's own engine)."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref, mine, X, y, ref_loss, my_loss = _pair()
    outs = []
    for layer, loss in ((ref, ref_loss), (mine, my_loss)):
        tn = layer.tensor_network
        k = dict(batch_size=40, lr=1.0, method="ridge_cholesky", eps=0.5)
        k.update(kw)
        order = k.pop("node_order", None)
        if order == "middle_two":
            k["node_order"] = tn.train_nodes[1:3]
        elif order == "tuple":
            k["node_order"] = (tn.train_nodes[:2], tn.train_nodes[2:])
        loose = k.pop("loose", False)
        stop = k.pop("stop_after", None)
        if stop is not None:
            calls = [0]

            def crit(calls=calls, stop=stop):
                calls[0] += 1
                return calls[0] >= stop
            k["convergence_criterion"] = crit
        outs.append(_trace(tn, X, y, loss, **k))
    (r_ret, r_ev, r_pred), (m_ret, m_ev, m_pred) = outs
    assert m_ret == r_ret
    assert [e[:3] for e in m_ev] == [e[:3] for e in r_ev]
    for a, b in zip(m_ev, r_ev):
        if a[0] == "loss":
            assert abs(a[3] - b[3]) <= (1e-2 if loose else 1e-7) * max(1.0, abs(b[3])), (a, b)
    if not loose:
        assert float((m_pred - r_pred).norm() / r_pred.norm()) < 1e-7


def test_unknown_method_raises_like_reference(monkeypatch):
    import fake_ops
    fake_ops.install(monkeypatch)
    ref, mine, X, y, ref_loss, my_loss = _pair()
    for layer, loss in ((ref, ref_loss), (mine, my_loss)):
        with pytest.raises(ValueError):
            layer.tensor_network.accumulating_swipe(X, y, loss, method="dogleg", eps=0.5)


def test_singular_system_returns_false_like_reference(monkeypatch):
    """A non-positive-definite local system (all-zero data, no ridge) ends the sweep with False (reference network.py:481-484)."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref, mine, X, y, ref_loss, my_loss = _pair()
    X0 = torch.zeros_like(X)
    for layer, loss in ((ref, ref_loss), (mine, my_loss)):
        assert layer.tensor_network.accumulating_swipe(X0, y, loss, method="cholesky", eps=0.0) is False


@pytest.mark.parametrize("loss_name", ["KLDivBregman", "SoftmaxSquaredLoss", "BinaryKLDivBregman", "UncertaintyAutogradLoss", "AutogradLoss"])
def test_reference_loss_objects_drive_the_engine(loss_name, monkeypatch):
    """Any object with the reference's ``forward(y_pred, y) -> (loss, d_loss, sqd_loss)`` contract (tensor/bregman.py) can be handed
    to the B200 engine: the per-sample Hessian is taken apart into rank-one terms on the host.  Side by side with the reference's own
    engine under the reference's class, the B200 engine under the mirrored class of tensor/bregman.py (values of the mirrored classes
    against the reference's: test_losses_vs_reference.py), same seed."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg = _import_reference()
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(3)
    N, F = 160, 3
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    if loss_name == "KLDivBregman":
        C = 2                                       # C logits + an appended zero logit against C + 1 class probabilities
        y = torch.tensor(np.eye(C + 1)[rng.integers(0, C + 1, N)])
        make = lambda b: b.KLDivBregman(w=1.0)
    elif loss_name == "SoftmaxSquaredLoss":
        C = 3
        y = torch.tensor(np.eye(C)[rng.integers(0, C, N)] * 0.9 + 0.1 / C)
        make = lambda b: b.SoftmaxSquaredLoss(w=1.0)
    elif loss_name == "BinaryKLDivBregman":
        C = 1
        y = torch.tensor(rng.integers(0, 2, size=(N, 1)).astype(np.float64))
        make = lambda b: b.BinaryKLDivBregman(w=1.0)
    elif loss_name == "UncertaintyAutogradLoss":
        C = 2                                       # (mean, pre-softplus std), default_train_uncertainty.py
        y = torch.tensor(np.tanh(X[:, 0].numpy()) + 0.1 * rng.normal(size=N))
        make = lambda b: b.UncertaintyAutogradLoss()
    else:
        C = 2
        y = torch.tensor(np.tanh(X[:, :2].numpy()) + 0.1 * rng.normal(size=(N, 2)))
        make = lambda b: b.AutogradLoss(torch.nn.HuberLoss(reduction="none", delta=0.5))
    kw = dict(batch_size=64, num_swipes=1, lr=0.5, method="ridge_cholesky", eps=2.0)
    out = []
    for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):       # the reference's loss class on its engine, the mirrored class on ours
        layer = mod.TensorTrainLayer(3, 3, F + 1, output_shape=C, constrict_bond=False, seed=11)
        tn = layer.tensor_network
        losses = []
        ok = tn.accumulating_swipe(X, y, make(breg), loss_callback=lambda NS, nd, l: losses.append(float(l)), **kw)
        out.append((ok, losses, tn.forward(X, to_tensor=True).detach()))
    (r_ok, r_l, r_p), (m_ok, m_l, m_p) = out
    assert m_ok == r_ok
    assert len(m_l) == len(r_l)
    if r_ok:
        for a, b in zip(m_l, r_l):
            assert abs(a - b) <= 1e-7 * max(1.0, abs(b)), (m_l, r_l)
        assert float((m_p - r_p).norm() / r_p.norm()) < 1e-7


@pytest.mark.parametrize("C,n,r,F", [(1, 3, 3, 3), (2, 4, 3, 2), (3, 2, 4, 4)])
def test_get_b_matches_the_reference(C, n, r, F, monkeypatch):
    """``get_b(node, grad)`` -- the right-hand side J^T grad the reference's matrix-free sweeps start from (network.py:259-291) --
    for every core of a train, from the right-hand-side kernel (stand-in here) against the reference's materialised-Jacobian einsum."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, _ = _import_reference()
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(C + 10 * n)
    N = 70
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    grad = torch.tensor(rng.normal(size=(N, C)))
    outs = []
    for mod in (ref_layers, tnb):
        tn = mod.TensorTrainLayer(n, r, F + 1, output_shape=C, constrict_bond=False, seed=4).tensor_network
        tn.forward(X, to_tensor=True)
        tn.set_input(X)
        outs.append([tn.get_b(nd, grad) for nd in tn.train_nodes])
    for a, b in zip(*outs):
        assert a.shape == b.shape
        assert float((a - b).norm() / a.norm()) < 1e-13
    with pytest.raises(RuntimeError):
        tnb.TensorTrainLayer(n, r, F + 1, output_shape=C, seed=4).tensor_network.get_b(None, grad)


def test_cpd_engine_runs_reference_built_graph(monkeypatch):
    """INTEGRATION.md route 1 for CPD: the reference's CPDLayer builds the graph, the B200 CPDNetwork runs it."""
    import fake_ops
    ref_layers, ref_breg = _import_reference()
    from tensornetworksfork_b200.tensor.cpd import CPDNetwork as FastCPD
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(1)
    N, F = 180, 4
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1].numpy()) + 0.1 * rng.normal(size=(N, 1)))
    kw = dict(batch_size=60, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)
    ref_layer = ref_layers.CPDLayer(3, 4, F + 1, output_shape=(1,), seed=5)
    ref_losses = []
    ref_layer.tensor_network.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: ref_losses.append(l), **kw)
    ref_pred = ref_layer.tensor_network.forward(X, to_tensor=True)
    layer2 = ref_layers.CPDLayer(3, 4, F + 1, output_shape=(1,), seed=5)
    old = layer2.tensor_network
    fast = FastCPD(old.input_nodes, old.main_nodes, old.train_nodes, output_labels=old.output_labels, sample_dim=old.sample_dim)
    layer2.set_tensor_network(fast)
    losses = []
    assert fast.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: losses.append(l), **kw)
    assert len(losses) == len(ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b))
    pred = fast.forward(X, to_tensor=True)
    assert float((pred.reshape(ref_pred.shape) - ref_pred).norm() / ref_pred.norm()) < 1e-8


def test_conv_engine_runs_reference_built_graph(monkeypatch):
    """INTEGRATION.md route 1 for the patch/pixel conv-TT: the reference's TensorConvolutionTrainLayer builds the graph (and grows
    it with its own grow_cart), ConvTrainNetwork recognises the columns and runs the sweeps the image scripts call."""
    import fake_ops
    ref_layers, ref_breg = _import_reference()
    from tensornetworksfork_b200.tensor.conv import ConvTrainNetwork
    from scipy.sparse.linalg import minres
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(2)
    N, Q, T = 90, 5, 4
    X = torch.tensor(rng.uniform(-1, 1, size=(N, Q, T)))
    X[:, -1, :] = 0.0
    X[:, :, -1] = 0.0
    X[:, -1, -1] = 1.0
    y = torch.tensor(np.eye(3)[rng.integers(0, 3, N)])
    ctor = dict(num_carriages=2, bond_dim=3, num_patches=Q, patch_pixels=T, output_shape=2, convolution_bond=2)
    outs = []
    for fast in (False, True):
        torch.manual_seed(4)
        layer = ref_layers.TensorConvolutionTrainLayer(**ctor)
        layer.grow_cart(3, 2)                                       # the reference's own growth; its graph lacks right_labels there
        if fast:
            old = layer.tensor_network
            layer.set_tensor_network(ConvTrainNetwork(old.input_nodes, old.main_nodes, old.train_nodes, output_labels=old.output_labels,
                                                      sample_dim=old.sample_dim))
        tn = layer.tensor_network
        loss = ref_breg.XEAutogradBregman(w=1.0)
        l1, l2 = [], []
        assert tn.accumulating_swipe(X, y, loss, batch_size=40, num_swipes=1, method="ridge_exact", eps=1.0, eps_decay=0.5,
                                     loss_callback=lambda NS, n, l: l1.append(float(l)))
        assert tn.scipy_swipe(X, y, loss, minres, batch_size=45, num_swipes=1, max_iter=4, tol=1e-8, loss_callback=lambda l: l2.append(float(l)))
        outs.append((l1, l2, tn.forward(X, to_tensor=True).detach()))
    (r1, r2, rp), (m1, m2, mp) = outs
    assert len(m1) == len(r1) and len(m2) == len(r2)
    for a, b in zip(m1, r1):
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (m1, r1)
    for a, b in zip(m2, r2):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (m2, r2)      # float32 Krylov recurrences on the host (network.py:918-926)
    assert float((mp - rp).norm() / rp.norm()) < 1e-3


def test_cumsum_engine_runs_reference_built_graph(monkeypatch):
    """INTEGRATION.md route 1 for the cum-sum train: the reference's CumSumLayer graph carries dense operator nodes between the
    inputs and the cores; CumSumNetwork reads the cores and inputs from it and applies the operator in closed form."""
    import fake_ops
    ref_layers, ref_breg = _import_reference()
    from tensornetworksfork_b200.tensor.cumsum import CumSumNetwork
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(3)
    N, F = 150, 3
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    y = torch.tensor(X[:, :1].numpy() * X[:, 1:2].numpy() + 0.05 * rng.normal(size=(N, 1)))
    kw = dict(batch_size=50, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)
    outs = []
    for fast in (False, True):
        torch.manual_seed(6)
        layer = ref_layers.CumSumLayer(3, 3, F + 1, output_shape=1, constrict_bond=False)
        if fast:
            old = layer.tensor_network
            layer.set_tensor_network(CumSumNetwork(old.input_nodes, old.main_nodes, old.train_nodes, output_labels=old.output_labels,
                                                   sample_dim=old.sample_dim))
        tn = layer.tensor_network
        losses = []
        assert tn.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: losses.append(float(l)), **kw)
        outs.append((losses, tn.forward(X, to_tensor=True).detach()))
    (rl, rp), (ml, mp) = outs
    assert len(ml) == len(rl)
    for a, b in zip(ml, rl):
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (ml, rl)
    assert float((mp.reshape(rp.shape) - rp).norm() / rp.norm()) < 1e-8


def test_linear_layer_engine_runs_reference_built_graph(monkeypatch):
    """INTEGRATION.md route 1 for TensorTrainLinearLayer (trainable projection W_k between input and core, layers.py:308-343)."""
    import fake_ops
    ref_layers, ref_breg = _import_reference()
    from tensornetworksfork_b200.tensor.network import TensorNetwork as FastTN
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(4)
    N, F = 170, 5
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1].numpy() - X[:, 1:2].numpy()) + 0.05 * rng.normal(size=(N, 1)))
    kw = dict(batch_size=60, num_swipes=2, lr=1.0, method="ridge_cholesky", eps=1.0, eps_decay=0.5)
    outs = []
    for fast in (False, True):
        layer = ref_layers.TensorTrainLinearLayer(3, 3, F + 1, 2, output_shape=1, constrict_bond=False, seed=8)
        if fast:
            old = layer.tensor_network
            layer.set_tensor_network(FastTN(old.input_nodes, old.main_nodes, old.train_nodes, output_labels=old.output_labels,
                                            sample_dim=old.sample_dim))
        tn = layer.tensor_network
        losses = []
        assert tn.accumulating_swipe(X, y, ref_breg.SquareBregFunction(), loss_callback=lambda NS, n, l: losses.append(float(l)), **kw)
        outs.append((losses, tn.forward(X, to_tensor=True).detach()))
    (rl, rp), (ml, mp) = outs
    assert len(ml) == len(rl)
    for a, b in zip(ml, rl):
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (ml, rl)
    assert float((mp.reshape(rp.shape) - rp).norm() / rp.norm()) < 1e-8
