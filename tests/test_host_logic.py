"""CPU-side checks: constructors reproduce the reference's initial state, the sweep schedule
reproduces its control flow, the C-ABI library loads and exports what include/tn_b200.h declares,
and the product path refuses to run without CUDA (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

import golden_util as gu

torch.set_default_dtype(torch.float64)

import tensornetworksfork_b200 as tnb  # noqa: E402
from tensornetworksfork_b200 import _lib  # noqa: E402
from tensornetworksfork_b200.tensor.layers import chain_bond_dims  # noqa: E402
from tensornetworksfork_b200.tensor.network import sweep_schedule, batch_mean_of_means  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONSTRUCT = {
    "tt_poly_reg": lambda: tnb.TensorTrainLayer(3, 4, 5, output_shape=1, constrict_bond=True, perturb=True, seed=42),
    "tt_poly5_full": lambda: tnb.TensorTrainLayer(5, 3, 4, output_shape=1, constrict_bond=False, perturb=False, seed=7),
    "tnml_poly_xe": lambda: tnb.TensorTrainLayer(5, 3, 3, output_shape=3, constrict_bond=True, perturb=False, seed=5),
    "tt_multi_square": lambda: tnb.TensorTrainLayer(3, 3, 4, output_shape=2, constrict_bond=False, perturb=False, seed=11),
    "cpd_reg": lambda: tnb.CPDLayer(4, 6, 4, output_shape=(1,), seed=42),
    "tt_exact_lr": lambda: tnb.TensorTrainLayer(2, 2, 3, output_shape=1, constrict_bond=False, perturb=False, seed=3),
}


@pytest.mark.parametrize("name", sorted(CONSTRUCT))
def test_constructors_reproduce_reference_init(name):
    """Same seed -> same shapes, labels order and values as the reference's layers (cores0 of the fixtures)."""
    fx = gu.load(name)
    layer = CONSTRUCT[name]()
    nodes = layer.tensor_network.train_nodes
    assert len(nodes) == len(fx["cores0"])
    for n, ref in zip(nodes, fx["cores0"]):
        assert tuple(n.shape) == ref.shape
        assert np.array_equal(n.tensor.numpy(), ref)


def test_bond_dims_table():
    assert chain_bond_dims(6, 8, 2) == [1, 2, 4, 8, 4, 2, 1]
    assert chain_bond_dims(7, 8, 2) == [1, 2, 4, 8, 8, 4, 2, 1]
    assert chain_bond_dims(5, 38, 29, constrict_bond=False) == [1, 38, 38, 38, 38, 1]
    assert chain_bond_dims(3, 6, 9, True, True) == [1, 6, 6, 1]
    assert chain_bond_dims(1, 8, 2) == [1, 1]
    d = chain_bond_dims(90, 24, 2)
    assert d[:7] == [1, 2, 4, 8, 16, 24, 24] and d[-6:] == [24, 16, 8, 4, 2, 1]


@pytest.mark.parametrize("name", [n for n in gu.names()])
def test_sweep_schedule_matches_reference_trace(name):
    fx = gu.load(name)
    meta = fx["meta"]
    n = len(fx["cores0"])
    cols = list(range(n)) if meta["kind"] == "tt" else list(range(n))
    direction = meta.get("direction", "l2r")
    first = cols if direction == "l2r" else cols[::-1]
    second = cols[::-1] if direction == "l2r" else cols
    sched = sweep_schedule(first, second, meta["num_swipes"], meta["eps"], meta.get("eps_decay"),
                           meta.get("skip_second", False), direction, meta.get("eps_per_node", False))
    got = [(NS, (first, second)[half][i], e) for NS, half, i, e in sched]
    want = [(u["NS"], u["k"], u["eps"]) for u in fx["updates"]]
    assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in want]
    for (_, _, e1), (_, _, e2) in zip(got, want):
        assert abs(e1 - e2) <= 1e-15 * max(1.0, abs(e2))


def test_schedule_single_node_updates_once():
    """With one train node the turn-around skip leaves a single update per call."""
    assert len(sweep_schedule([0], [0], 5, 1.0)) == 1


def test_schedule_eps_list_and_skip_second():
    s = sweep_schedule([0, 1, 2], [2, 1, 0], 2, [4.0, 3.0, 2.0, 1.0])
    assert [(ns, h, i) for ns, h, i, _ in s] == [(0, 0, 0), (0, 0, 1), (0, 0, 2), (1, 1, 1), (1, 1, 2), (2, 0, 1), (2, 0, 2),
                                                 (3, 1, 1), (3, 1, 2)]
    assert [e for *_, e in s] == [4.0, 4.0, 4.0, 3.0, 3.0, 2.0, 2.0, 1.0, 1.0]
    s2 = sweep_schedule([0, 1], [1, 0], 2, [4.0, 3.0], skip_second=True)
    assert [(ns, i) for ns, _, i, _ in s2] == [(0, 0), (0, 1), (1, 0), (1, 1)]


def test_schedule_short_eps_list_fails_where_the_reference_does():
    """An epsilon list shorter than the sweep is an IndexError in the reference only when the sweep gets to the missing entry
    (network.py:412,431): the schedule carries the error in place of the value, so a convergence criterion that ends the sweep
    earlier never sees it (TensorTrainRegressorEarlyStopping on a linear-projection train: N epsilons, 2N trained nodes)."""
    s = sweep_schedule([0, 1, 2, 3], [3, 2, 1, 0], 1, [4.0, 3.0], skip_second=True, eps_per_node=True)
    assert [e for *_, e in s][:2] == [4.0, 3.0]
    assert all(isinstance(e, IndexError) for *_, e in s[2:])
    s = sweep_schedule([0, 1], [1, 0], 2, [4.0, 3.0, 2.0])          # the fourth half-sweep has no epsilon
    assert [e for *_, e in s][:4] == [4.0, 4.0, 3.0, 2.0] and isinstance(s[-1][3], IndexError)


def test_early_stopping_regressor_on_a_linear_projection_train(monkeypatch):
    """tensor/module.py:573-582 passes one epsilon per core with eps_per_node=True; with linear_dim the train has two trained nodes per
    core, and the fit only works because early stopping ends the pass before the list runs out (found by a fuzz of the module.py
    classes against the reference: 750 configurations, this the only disagreement)."""
    import fake_ops
    import numpy as np
    from tensornetworksfork_b200.tensor.module import TensorTrainRegressorEarlyStopping
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(103)
    X = rng.uniform(-1, 1, size=(314, 5))
    y = np.tanh(X @ rng.normal(size=(5, 1))) + 0.3 * X[:, :1] * X[:, 1:2] + 0.05 * rng.normal(size=(314, 1))
    est = TensorTrainRegressorEarlyStopping(device="cpu", N=4, r=4, seed=22, constrict_bond=True, perturb=False, eps_start=0.1, eps_end=1.0,
                                            batch_size=500, method="ridge_cholesky", linear_dim=2, early_stopping=2)
    est.fit(X[:235], y[:235], X_val=X[235:], y_val=y[235:])
    assert np.isfinite(est.predict(X[235:])).all()
    # without a criterion the same call runs out of epsilons at the fifth trained node, as the reference does
    tn = est._model.tensor_network
    with pytest.raises(IndexError):
        tn.accumulating_swipe(torch.tensor(np.concatenate([X, np.ones((314, 1))], 1)), torch.tensor(y), est.bf, eps=[1.0] * 4, eps_per_node=True,
                              num_swipes=1, skip_second=True, method="ridge_cholesky", batch_size=500)


def test_batch_mean_of_means_overweights_short_batch():
    loss = torch.arange(10.0)
    got = batch_mean_of_means(loss, 4)
    want = (loss[:4].mean() + loss[4:8].mean() + loss[8:].mean()) / 3
    assert abs(float(got) - float(want)) < 1e-15
    assert abs(float(batch_mean_of_means(loss, -1)) - 4.5) < 1e-15


def test_library_exports_every_declared_symbol():
    """Every function include/tn_b200.h declares is exported and bound (no compute without a GPU)."""
    hdr = open(os.path.join(ROOT, "include", "tn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))
    assert lib.tn_version() >= 100


@pytest.mark.parametrize("rows,m", [(515345, (24, 2, 24)), (4177, (6, 9, 6)), (20640, (100, 1, 9)), (1000000, (38, 29, 38)), (131072, (38, 6, 38)),
                                    (60000, (10, 2, 10)), (300, (3, 3, 3))])
def test_gram_row_split_fills_whole_rounds(rows, m):
    """tn_gram_ksplit (fp64 Gram / right-hand side): every CTA of a launch does the same work and two are resident per SM, so a launch takes
    ceil(tiles * ks / slots) rounds -- the split must not leave the last round nearly empty (config 3's site once ran 600 CTAs on 296 slots),
    keeps at least 8 chunks of 32 rows per split, and needs no GPU to be asked (148 SMs assumed off-device)."""
    lib = _lib.load()
    ks = lib.tn_gram_ksplit(rows, m[0], m[1], m[2], 0)
    npair = lambda k: k * (k + 1) // 2
    nU, nV = npair(m[0]) * npair(m[1]), npair(m[2])
    tiles = -(-nU // 128) * -(-nV // 64)
    slots = 2 * 148
    assert 1 <= ks <= 1024
    assert ks == 1 or ks <= -(-rows // 256)
    assert lib.tn_gram_ksplit(rows, m[0], m[1], m[2], 3) == 1          # the tensor-core modes flush into M themselves
    if ks > 1:
        rounds = -(-tiles * ks // slots)
        if ks < -(-rows // 256):                      # not limited by the rows: the last round is (almost) full
            assert tiles * ks / (rounds * slots) > 0.85, (ks, tiles, rounds)
        # no split with fewer rounds per row of work is left on the table (2 % hysteresis, 64 rows of per-CTA overhead)
        cost = lambda k: -(-tiles * k // slots) * (-(-rows // k) + 64)
        assert cost(ks) <= min(cost(k) for k in range(1, min(1024, -(-rows // 256)) + 1) if tiles * k <= 8 * slots) * 1.021


def test_no_cpu_fallback():
    """On a CPU tensor the product path raises instead of computing."""
    layer = tnb.TensorTrainLayer(3, 2, 3, output_shape=1, seed=0)
    x = torch.rand(8, 3)
    with pytest.raises(RuntimeError):
        layer.tensor_network.forward(x, to_tensor=True)
    with pytest.raises(RuntimeError):
        layer.tensor_network.accumulating_swipe(x, torch.rand(8, 1), tnb.SquareBregFunction(), eps=1.0, method="ridge_cholesky")


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = "import sys; sys.path.insert(0, %r); import tensornetworksfork_b200, tensornetworksfork_b200.models; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT
    subprocess.check_call([sys.executable, "-c", code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tensornetworksfork_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_loss_closed_forms_match_rank1_terms():
    from tensornetworksfork_b200.tensor.bregman import hessian_terms
    torch.manual_seed(0)
    x = torch.randn(7, 3)
    y = torch.eye(4)[torch.randint(0, 4, (7,))]
    for lf, yy in ((tnb.XEAutogradBregman(w=0.7), y), (tnb.AutogradLoss(), torch.randn(7, 3)), (tnb.SquareBregFunction(), torch.randn(7, 3))):
        loss, g, H = lf.forward(x, yy)
        l2, g2, U, lam = hessian_terms(lf, x, yy)
        Hr = torch.einsum("st,stc,std->scd", lam, U, U)
        Hf = H.expand(7, 3, 3) if H.shape[-1] == 1 else H
        assert torch.allclose(Hr, Hf, atol=1e-14)
        assert torch.allclose(g, g2) and torch.allclose(loss, l2)


def test_kl_divergence_uses_the_soft_target_in_its_gradient():
    """KLDivBregman (reference tensor/bregman.py:100-146): loss from the arg-max label, gradient from the probability vector itself."""
    import importlib.util
    import sys
    import types
    torch.manual_seed(0)
    x = torch.randn(40, 3)
    y = torch.softmax(torch.randn(40, 4), dim=-1)          # soft labels
    loss, g, H = tnb.KLDivBregman(w=0.7).forward(x, y)
    z = torch.cat((0.7 * x, torch.zeros(40, 1)), dim=-1)
    p = torch.softmax(z, dim=-1)
    assert torch.allclose(g, 0.7 * (p - y)[:, :-1], atol=1e-14)
    assert torch.allclose(loss, -torch.log(p).gather(1, y.argmax(1, keepdim=True)).squeeze(1), atol=1e-13)
    t = tnb.KLDivBregman(w=0.7).rank1_terms(x, y)
    assert torch.allclose(t[1], g) and torch.allclose(torch.einsum("sv,svi,svj->sij", t[3], t[2], t[2]), H, atol=1e-13)
    if os.path.isdir("/root/reference/tensor"):
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))
        spec = importlib.util.spec_from_file_location("_ref_bregman", "/root/reference/tensor/bregman.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        rl, rg, rH = mod.KLDivBregman(w=0.7).forward(x, y)
        assert torch.allclose(rl, loss, atol=1e-13) and torch.allclose(rg, g, atol=1e-13) and torch.allclose(rH, H, atol=1e-13)


def test_qr_regauge_of_wide_cores_cpu(monkeypatch):
    """CPU twin of test_zz_gpu_late.py::test_qr_regauge_of_wide_cores_shrinks_the_bond_gpu on the stand-in kernels: the bonds of wide
    cores shrink exactly as the reference's reduced QR makes them (found by running the fuzz suite on 700 more seeds)."""
    import fake_ops
    import qr_wide_case as qw
    fake_ops.install(monkeypatch)
    mid, final, worst, drift = qw.run("cpu")
    assert mid == [(1, 2, 2), (2, 2, 4), (4, 2, 5), (5, 2)], mid
    assert final == [(1, 2, 2), (2, 2, 4), (4, 2, 2), (2, 2)], final
    assert worst < 1e-12 and drift < 1e-12, (worst, drift)
    mid, final, worst, drift = qw.run("cpu", sites=5, r=7, f=3, seed=11)
    assert worst < 1e-12 and drift < 1e-12, (worst, drift)


def test_stack_maintenance_calls_of_the_reference_are_accepted(monkeypatch):
    """recompute_all_stacks / left_update_stacks / right_update_stacks (reference network.py:73-77,152-172; called from outside the
    class by symmetric_operator.py:52 and cum_sum_operator.py:67): here they only forget cached environments, and a prediction after
    a core was changed in place is the one of the new cores either way."""
    import fake_ops
    import numpy as np
    import tensornetworksfork_b200 as tnb
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(2)
    X = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(40, 3)), np.ones((40, 1))], 1))
    y = torch.tensor(rng.normal(size=(40, 1)))
    tn = tnb.TensorTrainLayer(4, 3, 4, output_shape=1, constrict_bond=False, seed=1).tensor_network
    tn.accumulating_swipe(X, y, tnb.SquareBregFunction(), method="ridge_cholesky", eps=1.0, num_swipes=1)
    fresh = tnb.TensorTrainLayer(4, 3, 4, output_shape=1, constrict_bond=False, seed=1).tensor_network
    node = tn.main_nodes[1]
    node.tensor.mul_(1.25)
    tn.left_update_stacks(node)
    tn.right_update_stacks(node)
    for a, b in zip(fresh.main_nodes, tn.main_nodes):
        a.tensor = b.tensor.clone()
    assert float((tn.forward(X, to_tensor=True) - fresh.forward(X, to_tensor=True)).abs().max()) < 1e-13
    tn.recompute_all_stacks()
    assert float((tn.forward(X, to_tensor=True) - fresh.forward(X, to_tensor=True)).abs().max()) < 1e-13
    tnb.CPDLayer(3, 2, 4, output_shape=(1,), seed=1).tensor_network.recompute_all_stacks()
