"""Randomised side-by-side runs against the UNMODIFIED reference (build container only: needs /root/reference): seeded random
model shapes, losses and accumulating_swipe keyword combinations, the mirrored layer on the CPU stand-in kernels against the
reference layer on its own engine.  Well-conditioned settings only (ridge >= 0.3), so free-running sweeps stay comparable."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tensor")), reason="reference tree not mounted")
torch.set_default_dtype(torch.float64)


def _ref(module):
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.append(REF)
    return importlib.import_module(module)


def _draw(seed):
    rng = np.random.default_rng(1000 + seed)
    kind = ["tt", "tt", "tnml", "cpd"][seed % 4]
    n = int(rng.integers(1, 6)) if kind != "cpd" else int(rng.integers(2, 5))
    r = int(rng.integers(2, 5))
    F = int(rng.integers(2, 5))
    C = int(rng.choice([1, 1, 2, 3]))
    xe = C > 1 and bool(rng.integers(0, 2))
    N = int(rng.integers(40, 120))
    half_sweeps = int(rng.integers(1, 4))
    kw = dict(batch_size=int(rng.choice([-1, 16, 37, 1000])), num_swipes=half_sweeps, lr=float(rng.choice([1.0, 0.6])),
              method=str(rng.choice(["ridge_cholesky", "ridge_exact"])), skip_second=bool(rng.integers(0, 2)),
              direction=str(rng.choice(["l2r", "r2l"])))
    n_eps = half_sweeps if kw["skip_second"] else 2 * half_sweeps
    style = int(rng.integers(0, 3))
    if style == 0:
        kw["eps"] = float(rng.uniform(0.3, 2.0))
    elif style == 1:
        kw["eps"], kw["eps_decay"] = float(rng.uniform(1.0, 2.0)), float(rng.uniform(0.7, 0.95))
    else:
        kw["eps"] = [float(v) for v in rng.uniform(0.3, 2.0, size=n_eps)]
    if kind == "tt" and rng.integers(0, 3) == 0:
        kw["adaptive_step"], kw["max_norm"] = True, float(rng.uniform(1.0, 4.0))
    if kind == "tnml" and n > 1 and rng.integers(0, 2) == 0:
        kw["orthonormalize"] = True
    return dict(kind=kind, n=n, r=r, F=F, C=C, xe=xe, N=N, kw=kw, perturb=bool(kind == "tt" and C == 1 and rng.integers(0, 2)),
                constrict=bool(rng.integers(0, 2)), seed=int(rng.integers(0, 1000)))


@pytest.mark.parametrize("seed", list(range(32)) + [110])       # 110: an unconstricted TNML train whose QR re-gauge shrinks a bond
def test_random_configuration_side_by_side(seed, monkeypatch):
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg = _ref("tensor.layers"), _ref("tensor.bregman")
    import tensornetworksfork_b200 as tnb
    c = _draw(seed)
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(c["N"], c["F"]))
    if c["xe"]:
        y = torch.tensor(np.eye(c["C"] + 1)[rng.integers(0, c["C"] + 1, c["N"])])
    else:
        y = torch.tensor(np.tanh(X @ rng.normal(size=(c["F"], c["C"]))) + 0.1 * rng.normal(size=(c["N"], c["C"])))
    outs = []
    for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):
        if c["kind"] == "tnml":       # one site per feature, sin-cos map (models/tnml.py:11-16), a list of per-site inputs
            xs = [torch.tensor(np.stack([np.cos(0.5 * np.pi * X[:, j]), np.sin(0.5 * np.pi * X[:, j])], 1)) for j in range(c["F"])]
            layer = mod.TensorTrainLayer(c["F"], c["r"], 2, output_shape=c["C"], constrict_bond=c["constrict"], seed=c["seed"])
            x = xs
        else:
            x = torch.tensor(np.concatenate([X, np.ones((c["N"], 1))], 1))
            if c["kind"] == "cpd":
                layer = mod.CPDLayer(c["n"], c["r"], c["F"] + 1, output_shape=(c["C"],), seed=c["seed"])
            else:
                layer = mod.TensorTrainLayer(c["n"], c["r"], c["F"] + 1, output_shape=c["C"], constrict_bond=c["constrict"],
                                             perturb=c["perturb"], seed=c["seed"])
        loss = breg.XEAutogradBregman(w=1.0) if c["xe"] else breg.SquareBregFunction()
        tn = layer.tensor_network
        ev = []
        ret = tn.accumulating_swipe(x, y, loss, loss_callback=lambda NS, nd, l: ev.append((NS, tn.train_nodes.index(nd), float(l))), **c["kw"])
        outs.append((ret, ev, tn.forward(x, to_tensor=True).detach()))
    (r_ret, r_ev, r_p), (m_ret, m_ev, m_p) = outs
    assert m_ret == r_ret, c
    assert [e[:2] for e in m_ev] == [e[:2] for e in r_ev], c
    for a, b in zip(m_ev, r_ev):
        assert abs(a[2] - b[2]) <= 1e-6 * max(1.0, abs(b[2])), (c, a, b)
    assert float((m_p.reshape(r_p.shape) - r_p).norm() / max(float(r_p.norm()), 1e-12)) < 1e-6, c


def _draw_special(seed):
    rng = np.random.default_rng(5000 + seed)
    kind = ["type1", "cumsum", "linear", "conv", "conv_krylov", "type1_cpd"][seed % 6]
    c = dict(kind=kind, r=int(rng.integers(2, 4)), F=int(rng.integers(2, 5)), N=int(rng.integers(40, 100)), seed=int(rng.integers(0, 1000)),
             n=int(rng.integers(2, 4)), C=1, xe=False)
    if kind in ("conv", "conv_krylov", "linear"):
        c["C"] = int(rng.choice([1, 2]))
        c["xe"] = c["C"] > 1
    half = int(rng.integers(1, 3))
    c["kw"] = dict(batch_size=int(rng.choice([-1, 23, 64])), num_swipes=half, lr=1.0, method=str(rng.choice(["ridge_cholesky", "ridge_exact"])),
                   eps=float(rng.uniform(0.5, 2.0)), eps_decay=float(rng.uniform(0.7, 1.0)), direction=str(rng.choice(["l2r", "r2l"])))
    if kind == "conv_krylov":
        c["kw"] = dict(batch_size=int(rng.choice([16, 40])), num_swipes=1, lr=1.0, max_iter=int(rng.integers(2, 6)), tol=1e-10)
    return c


@pytest.mark.parametrize("seed", range(30))
def test_random_special_model_side_by_side(seed, monkeypatch):
    """Type-I sums (TT and CPD members), cum-sum train, linear-projection train, conv-TT (dense and matrix-free), random shapes."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg, ref_net = _ref("tensor.layers"), _ref("tensor.bregman"), _ref("tensor.network")
    import tensornetworksfork_b200 as tnb
    c = _draw_special(seed)
    rng = np.random.default_rng(seed)
    N, F, C = c["N"], c["F"], c["C"]
    if c["kind"].startswith("conv"):
        Q, T = F + 2, F + 1
        X = rng.uniform(-1, 1, size=(N, Q, T))
        X[:, -1, :] = 0.0
        X[:, :, -1] = 0.0
        X[:, -1, -1] = 1.0
        x = torch.tensor(X)
        feat = X[:, :-1, :-1].reshape(N, -1)
    else:
        X = rng.uniform(-1, 1, size=(N, F))
        x = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
        feat = X
    if c["xe"]:
        y = torch.tensor(np.eye(C + 1)[rng.integers(0, C + 1, N)])
    else:
        y = torch.tensor(np.tanh(feat @ rng.normal(size=(feat.shape[1], C)) / np.sqrt(feat.shape[1])) + 0.1 * rng.normal(size=(N, C)))
    outs = []
    for mod, breg, netmod in ((ref_layers, ref_breg, ref_net), (tnb, tnb, tnb)):
        torch.manual_seed(c["seed"])              # layers that do not apply their seed argument draw from the global generator
        k = c["kind"]
        if k == "type1":
            nets = [mod.TensorTrainLayer(i, bond_dim=c["r"], input_features=F if i != 1 else F + 1, output_shape=1, constrict_bond=True,
                                         perturb=True, seed=c["seed"] + i).tensor_network for i in range(1, c["n"] + 1)]
            layer = mod.TensorNetworkLayer(netmod.SumOfNetworks(nets, output_labels=nets[0].output_labels))
        elif k == "type1_cpd":
            nets = [mod.CPDLayer(i, c["r"], F if i != 1 else F + 1, output_shape=(1,), seed=c["seed"] + i).tensor_network
                    for i in range(1, c["n"] + 1)]
            layer = mod.TensorNetworkLayer(netmod.SumOfNetworks(nets, output_labels=nets[0].output_labels))
        elif k == "cumsum":
            layer = mod.CumSumLayer(c["n"], c["r"], F + 1, output_shape=1, constrict_bond=False)
        elif k == "linear":
            layer = mod.TensorTrainLinearLayer(c["n"], c["r"], F + 1, 2, output_shape=C, constrict_bond=False, seed=c["seed"])
        else:
            layer = mod.TensorConvolutionTrainLayer(num_carriages=c["n"], bond_dim=c["r"], num_patches=x.shape[1], patch_pixels=x.shape[2],
                                                    output_shape=C, convolution_bond=2)
        loss = breg.XEAutogradBregman(w=1.0) if c["xe"] else breg.SquareBregFunction()
        tn = layer.tensor_network
        ev = []
        if k == "conv_krylov":
            from scipy.sparse.linalg import minres
            ret = tn.scipy_swipe(x, y, loss, minres, loss_callback=lambda l: ev.append(float(l)), **c["kw"])
        else:
            ret = tn.accumulating_swipe(x, y, loss, loss_callback=lambda NS, nd, l: ev.append(float(l)), **c["kw"])
        outs.append((ret, ev, tn.forward(x, to_tensor=True).detach()))
    (r_ret, r_ev, r_p), (m_ret, m_ev, m_p) = outs
    tol = 1e-3 if c["kind"] == "conv_krylov" else 1e-6        # float32 Krylov recurrences on the host (network.py:918-926)
    assert m_ret == r_ret, c
    assert len(m_ev) == len(r_ev), c
    for a, b in zip(m_ev, r_ev):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (c, m_ev, r_ev)
    assert float((m_p.reshape(r_p.shape) - r_p).norm() / max(float(r_p.norm()), 1e-12)) < (1e-2 if c["kind"] == "conv_krylov" else 1e-6), c


@pytest.mark.parametrize("seed", [s for s in range(32) if s % 4 != 3])
def test_oracle_tracks_reference_on_random_configurations(seed):
    """The numpy oracle (the checker the GPU tests and smoke() use where no recording exists) against the live reference on the
    same random TT / TNML configurations: per-update loss trace and final prediction."""
    ref_layers, ref_breg = _ref("tensor.layers"), _ref("tensor.bregman")
    from oracle import tn_oracle as orc
    c = _draw(seed)
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(c["N"], c["F"]))
    if c["xe"]:
        y = np.eye(c["C"] + 1)[rng.integers(0, c["C"] + 1, c["N"])]
    else:
        y = np.tanh(X @ rng.normal(size=(c["F"], c["C"]))) + 0.1 * rng.normal(size=(c["N"], c["C"]))
    if c["kind"] == "tnml":
        xs = [np.stack([np.cos(0.5 * np.pi * X[:, j]), np.sin(0.5 * np.pi * X[:, j])], 1) for j in range(c["F"])]
        layer = ref_layers.TensorTrainLayer(c["F"], c["r"], 2, output_shape=c["C"], constrict_bond=c["constrict"], seed=c["seed"])
        x_np, x_t = xs, [torch.tensor(t) for t in xs]
    else:
        x_np = np.concatenate([X, np.ones((c["N"], 1))], 1)
        x_t = torch.tensor(x_np)
        layer = ref_layers.TensorTrainLayer(c["n"], c["r"], c["F"] + 1, output_shape=c["C"], constrict_bond=c["constrict"],
                                            perturb=c["perturb"], seed=c["seed"])
    tn = layer.tensor_network
    cores = [n.tensor.detach().numpy().copy() for n in tn.train_nodes]
    loss = ref_breg.XEAutogradBregman(w=1.0) if c["xe"] else ref_breg.SquareBregFunction()
    ref_ev = []
    ret = tn.accumulating_swipe(x_t, torch.tensor(y), loss, loss_callback=lambda NS, nd, l: ref_ev.append((NS, tn.train_nodes.index(nd), float(l))),
                                **c["kw"])
    trace = []
    ok = orc.accumulating_swipe(cores, x_np, y, loss="xe" if c["xe"] else "square", trace=trace, **c["kw"])
    assert ok == ret
    assert [(t["NS"], t["k"]) for t in trace] == [e[:2] for e in ref_ev], c
    for t, e in zip(trace, ref_ev):
        assert abs(t["loss"] - e[2]) <= 1e-6 * max(1.0, abs(e[2])), (c, t["loss"], e)
    pred = orc.forward(cores, x_np)
    ref_pred = tn.forward(x_t, to_tensor=True).detach().numpy()
    assert np.linalg.norm(pred.reshape(ref_pred.shape) - ref_pred) / max(np.linalg.norm(ref_pred), 1e-12) < 1e-6, c


@pytest.mark.parametrize("seed", range(24))
def test_random_matrix_free_sweeps_side_by_side(seed, monkeypatch):
    """lanczos_swipe (same torch seed -> same random start vectors in both engines) and scipy_swipe (cg / minres) on random TT,
    TNML, cum-sum and linear-projection models.  Few Krylov steps: the unregularised local systems are singular by gauge freedom and
    long recurrences amplify rounding differences."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg = _ref("tensor.layers"), _ref("tensor.bregman")
    import tensornetworksfork_b200 as tnb
    from scipy.sparse.linalg import cg, minres
    rng = np.random.default_rng(9000 + seed)
    kind = ["tt", "tnml", "cumsum", "linear"][seed % 4]
    solver = ["lanczos", "cg", "minres"][seed % 3]
    n, r, F = int(rng.integers(2, 5)), int(rng.integers(2, 4)), int(rng.integers(2, 5))
    C = 1 if kind == "cumsum" else int(rng.choice([1, 2]))
    xe = C > 1
    N = int(rng.integers(50, 110))
    X = rng.uniform(-1, 1, size=(N, F))
    y = torch.tensor(np.eye(C + 1)[rng.integers(0, C + 1, N)] if xe else np.tanh(X @ rng.normal(size=(F, C))) + 0.1 * rng.normal(size=(N, C)))
    kw = dict(batch_size=int(rng.choice([20, 64, 500])), num_swipes=int(rng.integers(1, 3)), lr=1.0, max_iter=int(rng.integers(2, 5)), tol=1e-10)
    mseed = int(rng.integers(0, 1000))
    outs = []
    for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):
        torch.manual_seed(mseed)
        if kind == "tnml":
            x = [torch.tensor(np.stack([np.cos(0.5 * np.pi * X[:, j]), np.sin(0.5 * np.pi * X[:, j])], 1)) for j in range(F)]
            layer = mod.TensorTrainLayer(F, r, 2, output_shape=C, constrict_bond=True, seed=mseed)
        else:
            x = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
            if kind == "cumsum":
                layer = mod.CumSumLayer(n, r, F + 1, output_shape=1, constrict_bond=False)
            elif kind == "linear":
                layer = mod.TensorTrainLinearLayer(n, r, F + 1, 2, output_shape=C, constrict_bond=False, seed=mseed)
            else:
                layer = mod.TensorTrainLayer(n, r, F + 1, output_shape=C, constrict_bond=False, seed=mseed)
        loss = breg.XEAutogradBregman(w=1.0) if xe else breg.SquareBregFunction()
        tn = layer.tensor_network
        ev = []
        torch.manual_seed(mseed + 1)              # the Lanczos start vectors
        if solver == "lanczos":
            ret = tn.lanczos_swipe(x, y, loss, loss_callback=lambda l: ev.append(float(l)), **kw)
        else:
            ret = tn.scipy_swipe(x, y, loss, {"cg": cg, "minres": minres}[solver], loss_callback=lambda l: ev.append(float(l)), **kw)
        outs.append((ret, ev, tn.forward(x, to_tensor=True).detach()))
    (r_ret, r_ev, r_p), (m_ret, m_ev, m_p) = outs
    assert m_ret == r_ret and len(m_ev) == len(r_ev)
    tol = 1e-6 if solver == "lanczos" else 2e-3          # SciPy path: float32 recurrences on the host (network.py:918-926)
    for a, b in zip(m_ev, r_ev):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (kind, solver, m_ev, r_ev)
    assert float((m_p.reshape(r_p.shape) - r_p).norm() / max(float(r_p.norm()), 1e-12)) < 10 * tol, (kind, solver)


@pytest.mark.parametrize("seed", range(18))
def test_random_driver_keywords_side_by_side(seed, monkeypatch):
    """The keywords the first fuzz does not draw: method='gradient' (its two halves differ, network.py:458-470 vs :558-584),
    eps_per_node, explicit node orders (list / tuple of two lists / subsets), convergence_criterion stops, timeout=0."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg = _ref("tensor.layers"), _ref("tensor.bregman")
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(7000 + seed)
    n, r, F = int(rng.integers(2, 6)), int(rng.integers(2, 4)), int(rng.integers(2, 5))
    C = int(rng.choice([1, 1, 2]))
    N = int(rng.integers(50, 120))
    X = rng.uniform(-1, 1, size=(N, F))
    x = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(np.tanh(X @ rng.normal(size=(F, C))) + 0.1 * rng.normal(size=(N, C)))
    mode = ["gradient", "per_node", "order_list", "order_tuple", "converge", "timeout"][seed % 6]
    kw = dict(batch_size=int(rng.choice([-1, 19, 50])), num_swipes=int(rng.integers(1, 3)), lr=1.0, method="ridge_cholesky",
              eps=float(rng.uniform(0.5, 2.0)), direction=str(rng.choice(["l2r", "r2l"])))
    if mode == "gradient":
        kw.update(method="gradient", lr=-float(rng.uniform(1e-4, 2e-3)), skip_second=bool(rng.integers(0, 2)))
    elif mode == "per_node":
        kw.update(eps=[float(v) for v in rng.uniform(0.3, 2.0, size=n)], eps_per_node=True, num_swipes=1, skip_second=bool(rng.integers(0, 2)))
    elif mode == "timeout":
        kw.update(timeout=0.0)
    mseed = int(rng.integers(0, 1000))
    pick = sorted(rng.choice(n, size=max(1, n - 1), replace=False).tolist())
    split = int(rng.integers(1, n)) if n > 1 else 1
    stop = int(rng.integers(1, 2 * n))
    outs = []
    for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):
        layer = mod.TensorTrainLayer(n, r, F + 1, output_shape=C, constrict_bond=bool(seed % 2), seed=mseed)
        tn = layer.tensor_network
        k = dict(kw)
        if mode == "order_list":
            k["node_order"] = [tn.train_nodes[i] for i in pick]
        elif mode == "order_tuple":
            k["node_order"] = (tn.train_nodes[:split], tn.train_nodes[split:][::-1] or tn.train_nodes[:1])
        elif mode == "converge":
            calls = [0]

            def crit(calls=calls):
                calls[0] += 1
                return calls[0] >= stop
            k["convergence_criterion"] = crit
        ev = []
        ret = tn.accumulating_swipe(x, y, breg.SquareBregFunction(), loss_callback=lambda NS, nd, l: ev.append((NS, tn.train_nodes.index(nd), float(l))),
                                    block_callback=lambda NS, nd: ev.append((NS, tn.train_nodes.index(nd), None)), **k)
        outs.append((ret, ev, tn.forward(x, to_tensor=True).detach()))
    (r_ret, r_ev, r_p), (m_ret, m_ev, m_p) = outs
    assert m_ret == r_ret, (mode, kw)
    assert [e[:2] for e in m_ev] == [e[:2] for e in r_ev], (mode, kw)
    for a, b in zip(m_ev, r_ev):
        assert (a[2] is None) == (b[2] is None)
        if a[2] is not None:
            assert abs(a[2] - b[2]) <= 1e-6 * max(1.0, abs(b[2])), (mode, kw, a, b)
    assert float((m_p.reshape(r_p.shape) - r_p).norm() / max(float(r_p.norm()), 1e-12)) < 1e-6, (mode, kw)


@pytest.mark.parametrize("seed", range(8))
def test_external_core_changes_between_and_inside_sweeps(seed, monkeypatch):
    """Cores changed behind the engine's back -- load_node_states in place / by replacement between sweeps on the SAME data
    object (cached environments must be dropped), and a callback that rescales a core in the middle of a sweep -- side by side
    with the reference, whose reset_stacks / set_input policy the engine replaces by identity + version stamps."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg = _ref("tensor.layers"), _ref("tensor.bregman")
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(8000 + seed)
    n, r, F = int(rng.integers(2, 5)), int(rng.integers(2, 4)), int(rng.integers(2, 5))
    N = int(rng.integers(50, 100))
    X = rng.uniform(-1, 1, size=(N, F))
    x = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    xv = torch.tensor(np.concatenate([rng.uniform(-1, 1, size=(30, F)), np.ones((30, 1))], 1))
    y = torch.tensor(np.tanh(X[:, :1]) + 0.1 * rng.normal(size=(N, 1)))
    kw = dict(batch_size=-1 if seed % 2 else 33, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=1.0)
    scale = float(rng.uniform(0.5, 1.5))
    which = int(rng.integers(0, n))
    outs = []
    for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):
        layer = mod.TensorTrainLayer(n, r, F + 1, output_shape=1, constrict_bond=False, seed=seed)
        tn = layer.tensor_network
        ev = []
        seen = [0]

        def meddle(NS, node, layer=layer, tn=tn, seen=seen):
            seen[0] += 1
            ev.append(float(layer(xv).abs().sum()))                       # a foreign forward in the middle of the sweep
            if seen[0] == 2:
                tn.train_nodes[which].tensor = tn.train_nodes[which].tensor * scale    # replaced from outside

        assert tn.accumulating_swipe(x, y, breg.SquareBregFunction(), block_callback=meddle, loss_callback=lambda NS, nd, l: ev.append(float(l)), **kw)
        snap = layer.node_states()
        assert tn.accumulating_swipe(x, y, breg.SquareBregFunction(), loss_callback=lambda NS, nd, l: ev.append(float(l)), **kw)
        layer.load_node_states(snap, set_value=bool(seed % 2))            # back to the snapshot, in place or by replacement
        assert tn.accumulating_swipe(x, y, breg.SquareBregFunction(), loss_callback=lambda NS, nd, l: ev.append(float(l)), **kw)
        outs.append((ev, layer(x).detach()))
    (r_ev, r_p), (m_ev, m_p) = outs
    assert len(m_ev) == len(r_ev)
    for a, b in zip(m_ev, r_ev):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (m_ev, r_ev)
    assert float((m_p.reshape(r_p.shape) - r_p).norm() / float(r_p.norm())) < 1e-6


@pytest.mark.parametrize("seed", range(8))
def test_random_growing_flows_side_by_side(seed, monkeypatch):
    """Topology changes between sweeps: the 2-site DMRG growth (grow_middle / block sweep / split_node, growing_DMRG.py:51-62)
    and the conv-TT growth (grow_cart, image_convolution_growing_MNIST.py:84-103), random sizes, against the reference."""
    import fake_ops
    fake_ops.install(monkeypatch)
    ref_layers, ref_breg = _ref("tensor.layers"), _ref("tensor.bregman")
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(6000 + seed)
    N = int(rng.integers(60, 110))
    outs = []
    if seed % 2 == 0:
        F, r = int(rng.integers(2, 5)), int(rng.integers(2, 5))
        X = rng.uniform(-1, 1, size=(N, F))
        x = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
        y = torch.tensor(np.tanh(X[:, :1] * X[:, 1:2]) + 0.1 * rng.normal(size=(N, 1)))
        grows = int(rng.integers(1, 3))
        for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):
            torch.manual_seed(seed)
            layer = mod.TensorTrainDMRGInfiLayer(r, F + 1, output_shape=1, constrict_bond=True)
            ev = []
            kw = dict(batch_size=-1, lr=1.0, method="ridge_cholesky", num_swipes=3, loss_callback=lambda NS, nd, l: ev.append(float(l)))
            assert layer.tensor_network.accumulating_swipe(x, y, breg.SquareBregFunction(), eps=1.0, **kw)
            for g in range(grows):
                layer.grow_middle()
                ev.append(float(layer(x).abs().sum()))
                assert layer.tensor_network.accumulating_swipe(x, y, breg.SquareBregFunction(), eps=0.3, **kw)
                node = layer.nodes[layer.num_carriages // 2]
                err = layer.split_node(node.dim_labels[:2], node.dim_labels[-2:], r, err=1e-8, is_last=g == grows - 1)
                ev.append(float(err))
            outs.append((ev, layer(x).detach()))
    else:
        Q, T, r, CB = int(rng.integers(3, 6)), int(rng.integers(3, 5)), int(rng.integers(2, 4)), int(rng.integers(1, 3))
        X = rng.uniform(-1, 1, size=(N, Q, T))
        X[:, -1, :] = 0.0
        X[:, :, -1] = 0.0
        X[:, -1, -1] = 1.0
        x = torch.tensor(X)
        y = torch.tensor(np.eye(3)[rng.integers(0, 3, N)])
        n0 = int(rng.integers(2, 4))          # (a single column under the cross-entropy loss raises inside the reference's own solve_system)
        for mod, breg in ((ref_layers, ref_breg), (tnb, tnb)):
            torch.manual_seed(seed)
            layer = mod.TensorConvolutionTrainLayer(num_carriages=n0, bond_dim=r, num_patches=Q, patch_pixels=T, output_shape=2, convolution_bond=CB)
            ev = []
            kw = dict(batch_size=40, lr=1.0, method="ridge_exact", eps=1.0, eps_decay=0.7, num_swipes=1)
            loss = breg.XEAutogradBregman(w=1.0)
            assert layer.tensor_network.accumulating_swipe(x, y, loss, loss_callback=lambda NS, nd, l: ev.append(float(l)), **kw)
            for g in range(2):
                layer.grow_cart(r, CB) if g == 0 else layer.grow_cart()
                ev.append(float(layer(x).abs().sum()))
                assert layer.tensor_network.accumulating_swipe(x, y, loss, direction="r2l" if g == 0 else "l2r",
                                                               loss_callback=lambda NS, nd, l: ev.append(float(l)), **kw)
            outs.append((ev, layer(x).detach()))
    (r_ev, r_p), (m_ev, m_p) = outs
    assert len(m_ev) == len(r_ev)
    for a, b in zip(m_ev, r_ev):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (m_ev, r_ev)
    assert float((m_p.reshape(r_p.shape) - r_p).norm() / float(r_p.norm())) < 1e-6
