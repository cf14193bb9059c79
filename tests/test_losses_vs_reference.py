"""Every loss oracle of the reference's tensor/bregman.py that works on real tensors, mirrored: values of (loss, gradient, Hessian) on
random inputs against the UNMODIFIED reference classes, the ``only_loss`` form, and the Bregman base class under a user potential.
Build container only (needs /root/reference); the closed forms are also checked against autograd without it."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
torch.set_default_dtype(torch.float64)
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tensor")), reason="reference tree not mounted")


def _ref():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.append(REF)
    return importlib.import_module("tensor.bregman")


def _phi(x):
    return (x ** 4).sum(-1, keepdim=True) + 0.5 * (x ** 2).sum(-1, keepdim=True)


CASES = {
    "square": (lambda m: m.SquareBregFunction(), "real"),
    "autograd_mse": (lambda m: m.AutogradLoss(), "real"),
    "autograd_huber": (lambda m: m.AutogradLoss(torch.nn.HuberLoss(reduction="none", delta=0.7)), "real"),
    "xe": (lambda m: m.XEAutogradBregman(w=1.3), "onehot"),
    "kl_soft": (lambda m: m.KLDivBregman(w=0.8), "prob"),
    "softmax_squared": (lambda m: m.SoftmaxSquaredLoss(w=1.7), "prob_full"),
    "binary_kl": (lambda m: m.BinaryKLDivBregman(w=0.9), "unit"),
    "autograd_bregman": (lambda m: m.AutogradBregman(_phi, d_phi_x_func=lambda x: 4 * x ** 3 + x), "real"),
    "uncertainty": (lambda m: m.UncertaintyAutogradLoss(), "nll"),
}


def _inputs(kind, seed, S=23, C=3):
    rng = np.random.default_rng(seed)
    x = torch.tensor(rng.normal(size=(S, C)))
    if kind == "real":
        y = torch.tensor(rng.normal(size=(S, C)))
    elif kind == "onehot":
        y = torch.tensor(np.eye(C + 1)[rng.integers(0, C + 1, S)])
    elif kind == "prob":
        p = rng.uniform(0.05, 1.0, size=(S, C + 1))
        y = torch.tensor(p / p.sum(1, keepdims=True))
    elif kind == "prob_full":
        p = rng.uniform(0.05, 1.0, size=(S, C))
        y = torch.tensor(p / p.sum(1, keepdims=True))
    elif kind == "unit":
        y = torch.tensor(rng.uniform(0, 1, size=(S, C)))
        y[0, 0], y[1, 1] = 0.0, 1.0            # the clamped ends
    else:
        x = torch.tensor(rng.normal(size=(S, 2)))
        y = torch.tensor(rng.normal(size=(S,)))
    return x, y


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_loss_oracle_matches_the_reference_class(name):
    import tensornetworksfork_b200 as tnb
    make, kind = CASES[name]
    for seed in range(3):
        x, y = _inputs(kind, seed)
        want = make(_ref()).forward(x.clone(), y.clone())
        got = make(tnb).forward(x.clone(), y.clone())
        assert len(got) == 3
        for g, w in zip(got, want):
            w = w.detach()
            assert tuple(g.shape) == tuple(w.shape), (name, g.shape, w.shape)
            assert float((g - w).abs().max()) <= 1e-12 * max(1.0, float(w.abs().max())), name
        if name not in ("square",):
            lo = make(tnb).forward(x.clone(), y.clone(), only_loss=True)
            assert float((lo - want[0].detach()).abs().max()) <= 1e-12 * max(1.0, float(want[0].abs().max()))


@needs_ref
def test_autograd_bregman_without_the_derivative_argument_fails_like_the_reference():
    import tensornetworksfork_b200 as tnb
    x, y = _inputs("real", 0)
    for mod in (_ref(), tnb):
        with pytest.raises(TypeError):
            mod.AutogradBregman(_phi).forward(x.clone(), y.clone())


def test_bregman_base_class_under_a_user_potential():
    """A subclass that only gives psi / d / dsq, as a user of the reference's BregFunction would write it: the divergence, its gradient
    and its Hessian against autograd."""
    import tensornetworksfork_b200 as tnb

    class Quartic(tnb.BregFunction):
        def psi(self, x):
            return (x ** 4).sum(-1)

        def d(self, x):
            return 4 * x ** 3

        def dsq(self, x):
            return torch.diag_embed(12 * x ** 2)

    x, y = _inputs("real", 4)
    loss, g, H = Quartic()(x, y)
    xo = x.clone().requires_grad_(True)
    ref = (xo ** 4).sum(-1) - (y ** 4).sum(-1) - (4 * y ** 3 * (xo - y)).sum(-1)
    gr = torch.autograd.grad(ref.sum(), xo)[0]
    assert float((loss - ref.detach()).abs().max()) < 1e-12 and float((g - gr).abs().max()) < 1e-12
    assert float((H - torch.diag_embed(12 * x ** 2)).abs().max()) == 0.0
    sq = tnb.SquareBregFunction()
    assert float((sq.psi(x) - (x ** 2).sum(-1)).abs().max()) == 0.0 and float((sq.d(x) - 2 * x).abs().max()) == 0.0
    assert tuple(sq.dsq(x).shape) == (23, 3, 1) and float(sq.prod(x, y).sub((x * y).sum(-1)).abs().max()) == 0.0


def test_closed_form_losses_against_autograd():
    """SoftmaxSquaredLoss gradient and BinaryKLDivBregman gradient / Hessian diagonal are the derivatives of their own losses."""
    import tensornetworksfork_b200 as tnb
    x, y = _inputs("prob_full", 7)
    f = tnb.SoftmaxSquaredLoss(w=1.7)
    xo = x.clone().requires_grad_(True)
    g = torch.autograd.grad(f.forward(xo, y, only_loss=True).sum(), xo)[0]
    assert float((f.forward(x, y)[1] - g).abs().max()) < 1e-12
    x, y = _inputs("unit", 8)
    y = y.clamp(0.05, 0.95)
    f = tnb.BinaryKLDivBregman(w=0.9)
    xo = x.clone().requires_grad_(True)
    g = torch.autograd.grad(f.forward(xo, y, only_loss=True).sum(), xo, create_graph=True)[0]
    h = torch.stack([torch.autograd.grad(g[:, i].sum(), xo, retain_graph=True)[0][:, i] for i in range(3)], 1)
    _, gg, hh = f.forward(x, y)
    assert float((gg - g.detach()).abs().max()) < 1e-12 and float((hh.squeeze(-1) - h).abs().max()) < 1e-12
