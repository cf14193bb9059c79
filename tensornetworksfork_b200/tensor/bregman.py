"""Loss oracles: ``forward(y_pred, y) -> (loss, d_loss[S,C], sqd_loss[S,C,C])``.

Same contract as the reference's tensor/bregman.py:9-14.  The three losses the named
configurations use are given in closed form (no autograd passes) and additionally expose
``rank1_terms``: the output Hessian written as ``sum_t lam[s,t] u[s,t,:] u[s,t,:]^T`` so the Gram
kernel can treat a C-class sample as V weighted rows.  Any other object with
the same ``forward`` works too; its Hessian is then eigendecomposed per sample.
"""
import torch
from torch import nn


class BregFunction(nn.Module):
    """Bregman divergence of a convex potential: D(x, y) = psi(x) - psi(y) - <psi'(y), x - y>, gradient psi'(x) - psi'(y), Hessian
    psi''(x) (reference tensor/bregman.py:5-32).  Subclasses give ``psi``, ``d`` (its gradient) and ``dsq`` (its Hessian, or its diagonal
    with a trailing axis of one); ``transform_forward`` may reshape the pair first and ``prod`` is the inner product over the last axis."""

    def transform_forward(self, x, y):
        return x, y

    def psi(self, x):
        raise NotImplementedError

    def d(self, x):
        raise NotImplementedError

    def dsq(self, x):
        raise NotImplementedError

    def prod(self, a, b):
        return (a * b).sum(dim=-1)

    def grad(self, x, y):
        return self.d(x) - self.d(y)

    def hess(self, x, y):
        return self.dsq(x)

    def forward(self, x, y):
        x, y = self.transform_forward(x, y)
        return self.psi(x) - self.psi(y) - self.prod(self.d(y), x - y), self.grad(x, y), self.hess(x, y)


class SquareBregFunction(BregFunction):
    """(x-y)^2 summed over outputs; g = 2(x-y); Hessian (S,C,1) filled with 2.

    Reference tensor/bregman.py:34-52.  With C > 1 the reference's Gram einsum broadcasts that
    Hessian over (c, c') -- an all-ones*2 block; kept bit for bit here:
    rank-1 term u = 1, lam = 2.
    """

    def transform_forward(self, x, y):
        return (x.flatten(start_dim=1) if x.ndim > 1 else x), (y.flatten(start_dim=1) if y.ndim > 1 else y)

    def psi(self, x):
        return torch.sum(x ** 2, dim=-1)

    def d(self, x):
        return 2 * x

    def dsq(self, x):
        return torch.full_like(x, 2).unsqueeze(-1)

    def forward(self, x, y):
        x, y = self.transform_forward(x, y)
        d = x - y
        loss = torch.sum(x ** 2, dim=-1) - torch.sum(y ** 2, dim=-1) - torch.sum(2 * y * d, dim=-1)
        return loss, 2 * x - 2 * y, torch.full_like(x, 2).unsqueeze(-1)

    def rank1_terms(self, x, y):
        loss, g, _ = self.forward(x, y)
        S, C = g.shape
        return loss, g, torch.ones((S, 1, C), dtype=g.dtype, device=g.device), torch.full((S, 1), 2.0, dtype=g.dtype, device=g.device)


class AutogradLoss(nn.Module):
    """Element-wise torch loss; gradient and Hessian by autograd (reference tensor/bregman.py:266-292).
    For the default MSE the closed form (g = 2(x-y), H = 2I) is used."""

    def __init__(self, loss_func=None):
        super().__init__()
        self._is_mse = loss_func is None or (isinstance(loss_func, nn.MSELoss) and loss_func.reduction == "none")
        self.loss_func = loss_func if loss_func is not None else nn.MSELoss(reduction="none")

    def forward(self, model_out, y_true, only_loss=False):
        if self._is_mse:
            d = model_out - y_true
            if only_loss:
                return d ** 2
            C = d.shape[-1]
            H = (2.0 * torch.eye(C, dtype=d.dtype, device=d.device)).expand(d.shape[0], C, C)
            return d ** 2, 2 * d, H
        with torch.enable_grad():
            xo = model_out.detach().clone().requires_grad_(True)
            loss = self.loss_func(xo, y_true)
            if only_loss:
                return loss.detach()
            g = torch.autograd.grad(loss.sum(), xo, create_graph=True)[0]
            rows = [torch.autograd.grad(g[..., i].sum(), xo, retain_graph=True, allow_unused=True)[0] for i in range(g.shape[-1])]
            rows = [r if r is not None else torch.zeros_like(xo) for r in rows]
            H = torch.stack(rows, dim=-2)
        return loss.detach(), g.detach(), H.detach()

    def rank1_terms(self, x, y):
        if not self._is_mse:
            return None
        loss, g, _ = self.forward(x, y)
        S, C = g.shape
        U = torch.eye(C, dtype=g.dtype, device=g.device).expand(S, C, C)
        return loss, g, U, torch.full((S, C), 2.0, dtype=g.dtype, device=g.device)


class XEAutogradBregman(nn.Module):
    """Softmax cross-entropy on logits [w*x, 0] (reference tensor/bregman.py:189-216), closed form
    of :100-146: p = softmax(z); g = w (p - y)[:-1]; H = w^2 (diag p - p p^T)[:-1,:-1]."""

    def __init__(self, w=1.0):
        super().__init__()
        self.w = w

    def _p(self, x):
        z = self.w * x
        z = torch.cat((z, torch.zeros_like(z[..., :1])), dim=-1)
        return torch.log_softmax(z, dim=-1)

    def forward(self, x, y, only_loss=False):
        logp = self._p(x)
        lab = y.argmax(dim=-1)
        loss = -logp.gather(-1, lab.unsqueeze(-1)).squeeze(-1)
        if only_loss:
            return loss
        p = logp.exp()
        yoh = torch.zeros_like(p).scatter_(-1, lab.unsqueeze(-1), 1.0)
        g = self.w * (p - yoh)[..., :-1]
        H = (self.w ** 2) * (torch.diag_embed(p) - p.unsqueeze(-1) * p.unsqueeze(-2))[..., :-1, :-1]
        return loss, g, H

    def rank1_terms(self, x, y):
        logp = self._p(x)
        lab = y.argmax(dim=-1)
        loss = -logp.gather(-1, lab.unsqueeze(-1)).squeeze(-1)
        p = logp.exp()
        yoh = torch.zeros_like(p).scatter_(-1, lab.unsqueeze(-1), 1.0)
        g = self.w * (p - yoh)[..., :-1]
        S, C = g.shape
        pc = p[..., :-1]
        U = torch.cat([torch.eye(C, dtype=g.dtype, device=g.device).expand(S, C, C), pc.unsqueeze(1)], dim=1)
        lam = torch.cat([(self.w ** 2) * pc, torch.full((S, 1), -(self.w ** 2), dtype=g.dtype, device=g.device)], dim=1)
        return loss, g, U, lam


class KLDivBregman(XEAutogradBregman):
    """KL divergence to a target probability vector on logits [w*x, 0] (reference tensor/bregman.py:100-146): the reported loss is
    the cross-entropy of the arg-max label (:127), but the gradient uses the target vector ITSELF, g = w (p - y)[:-1] (:143) -- soft
    labels pull towards y, not towards its arg-max.  The Hessian is that of XEAutogradBregman."""

    def __init__(self, w=1.0, grad_clip=1e3):
        super().__init__(w=w)
        self.grad_clip = grad_clip          # kept for the signature; the reference's clipped expression is commented out (:134-139)

    def forward(self, x, y, only_loss=False):
        out = super().forward(x, y, only_loss=only_loss)
        if only_loss:
            return out
        loss, _, H = out
        return loss, self.w * (self._p(x).exp() - y)[..., :-1], H

    def rank1_terms(self, x, y):
        loss, _, U, lam = super().rank1_terms(x, y)
        return loss, self.w * (self._p(x).exp() - y)[..., :-1], U, lam


def _autograd_terms(loss_of, x):
    """(loss, gradient, Hessian rows) of an element-wise loss by autograd, as the reference's autograd losses compute them
    (tensor/bregman.py:201-214): one backward pass for the gradient, one per output for the Hessian."""
    with torch.enable_grad():
        xo = x.detach().clone().requires_grad_(True)
        loss = loss_of(xo)
        g = torch.autograd.grad(loss.sum(), xo, create_graph=True)[0]
        rows = [torch.autograd.grad(g[..., i].sum(), xo, retain_graph=True, allow_unused=True)[0] for i in range(g.shape[-1])]
        rows = [r if r is not None else torch.zeros_like(xo) for r in rows]
        H = torch.stack(rows, dim=-2)
    return loss.detach(), g.detach(), H.detach()


class SoftmaxSquaredLoss(nn.Module):
    """0.5 |softmax(w x) - y|^2 (kept with a trailing axis of one) with the Gauss-Newton Hessian w^2 J J^T of the softmax Jacobian
    J = diag(s) - s s^T, the third-order term dropped (reference tensor/bregman.py:68-98)."""

    def __init__(self, w=1.0):
        super().__init__()
        self.w = w

    def forward(self, x, y, only_loss=False):
        s = torch.softmax(self.w * x, dim=-1)
        r = s - y
        loss = 0.5 * (r * r).sum(dim=-1, keepdim=True)
        if only_loss:
            return loss
        J = torch.diag_embed(s) - s.unsqueeze(-1) * s.unsqueeze(-2)          # symmetric
        return loss, self.w * (J @ r.unsqueeze(-1)).squeeze(-1), (self.w ** 2) * (J @ J.transpose(-1, -2))


class BinaryKLDivBregman(BregFunction):
    """Element-wise KL divergence between Bernoulli(y) and Bernoulli(sigmoid(w x)), both clamped to [eps, 1 - eps]; gradient
    w (s - y), diagonal Hessian w^2 s (1 - s) with a trailing axis of one (reference tensor/bregman.py:148-187)."""

    def __init__(self, w=1.0):
        super().__init__()
        self.w = w

    def forward(self, x, y, only_loss=False, eps=1e-12):
        s = torch.sigmoid(self.w * x).clamp(min=eps, max=1 - eps)
        y = y.clamp(min=eps, max=1 - eps)
        kl = y * torch.log(y / s) + (1 - y) * torch.log((1 - y) / (1 - s))      # after the clamp neither branch of the reference's where() is empty
        if only_loss:
            return kl
        return kl, self.w * (s - y), ((self.w ** 2) * s * (1 - s)).unsqueeze(-1)


class AutogradBregman(BregFunction):
    """Bregman divergence of a user potential ``phi_func`` with gradient and Hessian by autograd (reference tensor/bregman.py:218-263).
    As there, ``d_phi_x_func`` only has to be given -- the derivative of phi is taken by autograd when it is, and the call fails with
    the reference's ``TypeError`` when it is not (its branches are the wrong way round, :238-243)."""

    def __init__(self, phi_func, forward_transform=None, d_phi_x_func=None):
        super().__init__()
        self.phi_func = phi_func
        self._transform_forward = forward_transform
        self._d_phi_x_func = d_phi_x_func

    def transform_forward(self, x, y):
        if self._transform_forward is not None:
            x, y = self._transform_forward(x, y)
        return x, y

    def _divergence(self, x, y):
        x, y = self.transform_forward(x, y)
        phi_x = self.phi_func(x)
        if self._d_phi_x_func is None:
            raise TypeError("'NoneType' object is not callable")
        d_phi_x = torch.autograd.grad(phi_x.sum(), x, create_graph=True)[0]
        return self.phi_func(y) - phi_x - (d_phi_x * (y - x)).sum(-1, keepdim=True)

    def forward(self, x, y, only_loss=False):
        if only_loss:
            with torch.enable_grad():
                return self._divergence(x.detach().clone().requires_grad_(True), y).detach()
        return _autograd_terms(lambda xo: self._divergence(xo, y), x)


class UncertaintyAutogradLoss(nn.Module):
    """Negative log-likelihood of y under Normal(mean = y_pred[..., 0], std = softplus(y_pred[..., 1])); gradient and 2 x 2 Hessian per
    sample by autograd (reference tensor/bregman.py:296-327)."""

    def forward(self, y_pred, y_true, only_loss=False):
        def nll(p):
            return -torch.distributions.Normal(loc=p[..., 0], scale=torch.nn.functional.softplus(p[..., 1])).log_prob(y_true)
        if only_loss:
            return nll(y_pred.detach())
        return _autograd_terms(nll, y_pred)


def hessian_terms(loss_fn, y_pred, y):
    """(loss, g[S,C], U[S,V,C], lam[S,V]) for any loss object.  Closed forms when the loss offers
    them; otherwise a per-sample eigendecomposition of the symmetrised Hessian it returned."""
    if hasattr(loss_fn, "rank1_terms"):
        t = loss_fn.rank1_terms(y_pred, y)
        if t is not None:
            return t
    loss, g, H = loss_fn.forward(y_pred, y)
    S, C = g.shape
    if H.dim() == 2:
        H = H.unsqueeze(-1)
    H = H.expand(S, C, C) if H.shape[-1] != C or H.shape[-2] != C else H
    if C == 1:
        return loss, g, torch.ones((S, 1, 1), dtype=g.dtype, device=g.device), H.reshape(S, 1)
    Hs = 0.5 * (H + H.transpose(-1, -2))
    lam, vec = torch.linalg.eigh(Hs)
    return loss, g, vec.transpose(-1, -2).contiguous(), lam.contiguous()
