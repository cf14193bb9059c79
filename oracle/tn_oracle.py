"""CPU oracle for the per-site Gauss-Newton / ALS sweep.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithm the reference
(niccogc/TensorNetworksFork, mounted at /root/reference while building) runs in
``tensor/network.py``.  It exists so the CUDA path has something to be checked
against on a box where the reference itself is absent.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; the product package never does.

Pinning: the reference stores no golden vectors and has no tests,
so this oracle is pinned against outputs of the reference itself run in the build
container -- ``tests/golden/make_golden.py`` generates them, ``tests/golden/*.npz``
holds them, ``tests/test_oracle_golden.py`` compares.  The arithmetic underneath
the reference is PyTorch 2.11 / SciPy 1.18 here (pins in the reference: torch
2.5.1, scipy 1.15.2, environment.yml:11-12).

Conventions
-----------
A tensor-train chain is a list of cores in the reference's *squeezed* layouts
(tensor/layers.py:77-97, node.py:135-147): site 1 ``(C, f, r)``, middle sites
``(r_l, f, r_r)``, last site ``(r, f)``; a single-site chain is ``(C, f)``.
Internally every core is viewed as ``(r_l, c, f, r_r)`` with ``c = C`` on the
site that owns the output leg and 1 elsewhere.  Environments are ``(S, c, r)``.
The parameter order of a core (rows/cols of ``A``, entries of ``b``) is the
row-major order of the squeezed layout, which equals ``(a, c, p, b)`` order.
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------- feature maps


def fbasis(X):
    """cos/sin(pi x / 2) per feature.  Reference: models/tnml.py:11-16."""
    X = np.asarray(X, dtype=np.float64)
    return [np.stack([np.cos(0.5 * math.pi * X[:, i]), np.sin(0.5 * math.pi * X[:, i])], axis=-1)
            for i in range(X.shape[1])]


def polynomial_basis(X, degree=3):
    """[x^0 .. x^degree] per feature.  Reference: models/tnml.py:18-23."""
    X = np.asarray(X, dtype=np.float64)
    return [np.stack([X[:, i] ** d for d in range(degree + 1)], axis=-1) for i in range(X.shape[1])]


def append_bias(X):
    """[X | 1]: bias column appended last.  Reference: models/tensor_train.py:223."""
    X = np.asarray(X, dtype=np.float64)
    return np.concatenate([X, np.ones((X.shape[0], 1))], axis=1)


# --------------------------------------------------------------------------- losses


def loss_square(pred, y):
    """SquareBregFunction: loss=(x-y)^2 summed over C, g=2(x-y), H=2 broadcast
    over (c,c').  Reference: tensor/bregman.py:9-14,34-52 (the Hessian there is
    (S,C,1) and the Gram einsum broadcasts it to all-ones*2 over (c,c'),
    network.py:188,212 -- SURVEY.md §7.3 item 4)."""
    pred = pred.reshape(pred.shape[0], -1)
    y = y.reshape(y.shape[0], -1)
    d = pred - y
    loss = (pred ** 2).sum(-1) - (y ** 2).sum(-1) - (2 * y * d).sum(-1)
    C = pred.shape[1]
    H = np.full((pred.shape[0], C, C), 2.0)
    return loss, 2 * pred - 2 * y, H


def loss_mse_autograd(pred, y):
    """AutogradLoss(MSELoss(reduction='none')): loss (S,C), g=2(x-y), H=2*I.
    Reference: tensor/bregman.py:266-292."""
    d = pred - y
    C = pred.shape[1]
    H = np.broadcast_to(2.0 * np.eye(C), (pred.shape[0], C, C)).copy()
    return d ** 2, 2 * d, H


def loss_xe(pred, y, w=1.0):
    """XEAutogradBregman closed form: p = softmax([w x, 0]); loss = CE;
    g = w (p - y)[:-1]; H = w^2 (diag p - p p^T)[:-1,:-1].
    Reference: tensor/bregman.py:189-216 (autograd), 100-146 (closed form)."""
    z = np.concatenate([w * pred, np.zeros((pred.shape[0], 1))], axis=1)
    z = z - z.max(axis=1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
    p = np.exp(logp)
    lab = y.argmax(axis=1)
    loss = -logp[np.arange(len(lab)), lab]
    yoh = np.zeros_like(p)
    yoh[np.arange(len(lab)), lab] = 1.0
    g = w * (p - yoh)[:, :-1]
    Hfull = w * w * (p[:, :, None] * np.eye(p.shape[1])[None] - p[:, :, None] * p[:, None, :])
    return loss, g, Hfull[:, :-1, :-1]


LOSSES = {"square": loss_square, "mse": loss_mse_autograd, "xe": loss_xe}

# --------------------------------------------------------------------------- chain helpers


def canon(core, k, n):
    """Squeezed reference layout -> (r_l, c, f, r_r)."""
    c = np.asarray(core, dtype=np.float64)
    if n == 1:
        return c.reshape(1, c.shape[0], c.shape[1], 1)
    if k == 0:
        return c.reshape(1, c.shape[0], c.shape[1], c.shape[2])
    if k == n - 1:
        return c.reshape(c.shape[0], 1, c.shape[1], 1)
    return c.reshape(c.shape[0], 1, c.shape[1], c.shape[2])


def site_inputs(x, n):
    """One (N,f) matrix shared by every site, or one per site.
    Reference: tensor/network.py:329-341."""
    if isinstance(x, (list, tuple)):
        return [np.asarray(t, dtype=np.float64) for t in x]
    x = np.asarray(x, dtype=np.float64)
    return [x] * n


def _absorb(env, phi, G4, left=True):
    """One environment step.  Reference: tensor/network.py:55-71,152-172
    (pairwise einsums through node.py:28-74)."""
    if left:
        out = np.einsum("sxa,sp,aypb->sxyb", env, phi, G4, optimize=True)
    else:
        out = np.einsum("sxb,sp,aypb->sxya", env, phi, G4, optimize=True)
    S, x_, y_, r = out.shape
    assert x_ == 1 or y_ == 1
    return out.reshape(S, x_ * y_, r)


def left_envs(cores, phis):
    """L_k for k = 0..n-1 (L_k includes site k).  network.py:55-71."""
    n = len(cores)
    S = phis[0].shape[0]
    env = np.ones((S, 1, 1))
    out = []
    for k in range(n):
        env = _absorb(env, phis[k], canon(cores[k], k, n), True)
        out.append(env)
    return out


def right_envs(cores, phis):
    """R_k for k = 0..n-1 (R_k includes site k).  network.py:55-71."""
    n = len(cores)
    S = phis[0].shape[0]
    env = np.ones((S, 1, 1))
    out = [None] * n
    for k in reversed(range(n)):
        env = _absorb(env, phis[k], canon(cores[k], k, n), False)
        out[k] = env
    return out


def forward(cores, x):
    """Prediction (S, C).  network.py:115-137."""
    phis = site_inputs(x, len(cores))
    return left_envs(cores, phis)[-1][:, :, 0]


def jacobian(cores, phis, k, L=None, R=None):
    """Local Jacobian J[s, c, P] of the prediction w.r.t. core k, P in the
    squeezed row-major parameter order.  network.py:101-113, 183
    (expand_labels: if core k owns the output leg, J is block-diagonal in c)."""
    n = len(cores)
    S = phis[0].shape[0]
    G4 = canon(cores[k], k, n)
    rl, ck, f, rr = G4.shape
    if L is None:
        L = left_envs(cores, phis)[k - 1] if k > 0 else np.ones((S, 1, 1))
    if R is None:
        R = right_envs(cores, phis)[k + 1] if k < n - 1 else np.ones((S, 1, 1))
    cl, cr = L.shape[1], R.shape[1]
    C = max(cl, ck, cr)
    J = np.zeros((S, C, rl, ck, f, rr))
    if ck > 1:
        base = np.einsum("sa,sp,sb->sapb", L[:, 0], phis[k], R[:, 0])
        for c in range(C):
            J[:, c, :, c] = base
    else:
        Lc = np.broadcast_to(L, (S, C, rl)) if cl == 1 else L
        Rc = np.broadcast_to(R, (S, C, rr)) if cr == 1 else R
        J[:, :, :, 0] = np.einsum("sca,sp,scb->scapb", Lc, phis[k], Rc)
    return J.reshape(S, C, rl * ck * f * rr)


def gram(J, g, H):
    """A = sum_s J^T H J, b = sum_s J^T g.  network.py:174-217 (contracted in
    the (J H)-first order, i.e. the opt_einsum path of SURVEY.md §8d baseline B;
    the sums are the same numbers)."""
    S, C, P = J.shape
    if C == 1:
        Jw = J[:, 0] * H[:, 0, 0][:, None]
        A = Jw.T @ J[:, 0]
        b = J[:, 0].T @ g[:, 0]
    else:
        JH = np.einsum("scd,sdP->scP", H, J)
        A = JH.reshape(S * C, P).T @ J.reshape(S * C, P)
        b = J.reshape(S * C, P).T @ g.reshape(S * C)
    return A, b


try:                                         # imported here, not inside the solve: the import costs ~1 s once
    from scipy.linalg import solve_triangular as _solve_triangular
except ImportError:
    _solve_triangular = None


def _cholesky_solve(A_f, rhs):
    """Cholesky factorisation + two TRIANGULAR substitutions (torch.linalg.cholesky + torch.cholesky_solve, network.py:313-315).
    The factorisation stays in numpy's BLAS (the Gram GEMM just ran there; a second BLAS pool -- SciPy's or torch's -- would
    fight it for the cores and cost 5-10x); the O(P^2) substitutions go through SciPy's dtrsv, with a general-solve fallback that
    gives the same numbers.  Not positive definite -> np.linalg.LinAlgError."""
    Lc = np.linalg.cholesky(A_f)
    if _solve_triangular is None:
        return np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
    y = _solve_triangular(Lc, rhs, lower=True, check_finite=False)
    return _solve_triangular(Lc, y, lower=True, trans="T", check_finite=False)


def solve_system(A, b, theta, method="exact", eps=0.0):
    """Scaled / ridge-regularised local solve.  network.py:293-327.
    Raises np.linalg.LinAlgError where torch raises LinAlgError."""
    P = b.size
    A_f = np.array(A, dtype=np.float64).reshape(P, P)
    b_f = np.array(b, dtype=np.float64).reshape(P)
    scale = np.abs(np.diag(A_f)).mean()
    if scale == 0:
        scale = 1.0
    A_f = A_f / scale
    b_f = b_f / scale
    m = method.lower()
    if m == "exact":
        x = np.linalg.solve(A_f, -b_f)
    elif m == "ridge_exact":
        A_f = A_f + (2 * eps) * np.eye(P)
        b_f = b_f + (2 * eps) * np.asarray(theta).reshape(P)
        x = np.linalg.solve(A_f, -b_f)
    elif m.startswith("ridge_cholesky") or m == "cholesky":
        if m != "cholesky":
            A_f = A_f + (2 * eps) * np.eye(P)
            b_f = b_f + (2 * eps) * np.asarray(theta).reshape(P)
        x = _cholesky_solve(A_f, -b_f)
    elif m == "gradient":
        x = -np.asarray(b, dtype=np.float64).reshape(P)
    else:
        raise ValueError(f"Unknown method: {method}")
    return x.reshape(np.shape(b))


def update_node(theta, step, lr=1.0, adaptive_step=False, max_norm=None):
    """theta <- theta + lr*step with optional shrink / projection.  node.py:178-203."""
    if adaptive_step:
        sn, pn = np.linalg.norm(step), np.linalg.norm(theta)
        if sn > pn:
            step = step * (pn / sn)
    new = theta + lr * step
    if max_norm is not None:
        cn = np.linalg.norm(new)
        if cn > max_norm:
            new = new * (max_norm / cn)
    return new


def _householder_qr(M):
    """Reduced QR with LAPACK's sign convention (what torch.linalg.qr returns)."""
    return np.linalg.qr(M, mode="reduced")


def orthonormalize_left(cores, k):
    """QR of core k as (r_l*c*f, r_r); R pushed into core k+1.  network.py:625-660."""
    n = len(cores)
    if k >= n - 1:
        return
    c = cores[k]
    M = c.reshape(-1, c.shape[-1])
    Q, Rm = _householder_qr(M)
    cores[k] = Q.reshape(c.shape[:-1] + (Q.shape[-1],))
    nxt = cores[k + 1]
    cores[k + 1] = np.tensordot(Rm, nxt, axes=(1, 0))


def orthonormalize_right(cores, k):
    """RQ (QR of the doubly flipped matrix) of core k as (c*f*r_r, r_l); factor
    pushed into core k-1.  network.py:662-707."""
    if k <= 0:
        return
    c = cores[k]
    perm = tuple(range(1, c.ndim)) + (0,)
    Ap = np.transpose(c, perm)
    shp = Ap.shape
    M = Ap.reshape(-1, shp[-1])
    Qr, Rr = _householder_qr(M[::-1, ::-1])
    Rm = Rr.T[::-1, ::-1]
    Q = Qr[::-1, ::-1]
    Q = Q.reshape(shp[:-1] + (Q.shape[-1],))
    inv = np.argsort(perm)
    cores[k] = np.transpose(Q, inv)
    prv = cores[k - 1]
    cores[k - 1] = np.einsum("ji,...j->...i", Rm, prv)


# --------------------------------------------------------------------------- one site update


def site_update(cores, x, y, k, loss="square", batch_size=-1, method="ridge_cholesky", eps=0.0, lr=1.0,
                loss_kwargs=None, adaptive_step=False, max_norm=None, apply=True):
    """Everything accumulating_swipe does for ONE node (network.py:438-486):
    minibatch accumulation of A, b and the mean-of-batch-means loss, solve, update.
    Returns a dict with A, b, step, loss, new core."""
    n = len(cores)
    phis_all = site_inputs(x, n)
    N = phis_all[0].shape[0]
    bs = N if batch_size <= 0 else batch_size
    nb = (N + bs - 1) // bs
    lf = LOSSES[loss] if isinstance(loss, str) else loss
    kw = loss_kwargs or {}
    A_out = b_out = None
    tot = 0.0
    y = np.asarray(y, dtype=np.float64)
    for bi in range(nb):
        sl = slice(bi * bs, (bi + 1) * bs)
        phis = [p[sl] for p in phis_all]
        Ls = left_envs(cores, phis)
        Rs = right_envs(cores, phis)
        pred = Ls[-1][:, :, 0]
        lo, g, H = lf(pred, y[sl].reshape(pred.shape[0], -1), **kw)
        S = pred.shape[0]
        L = Ls[k - 1] if k > 0 else np.ones((S, 1, 1))
        R = Rs[k + 1] if k < n - 1 else np.ones((S, 1, 1))
        J = jacobian(cores, phis, k, L, R)
        A, b = gram(J, g, H)
        A_out = A if A_out is None else A_out + A
        b_out = b if b_out is None else b_out + b
        tot += float(np.mean(lo))
    theta = np.asarray(cores[k], dtype=np.float64)
    step = solve_system(A_out, b_out, theta.reshape(-1), method=method, eps=eps).reshape(theta.shape)
    new = update_node(theta, step, lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
    if apply:
        cores[k] = new
    return {"A": A_out, "b": b_out, "step": step, "loss": tot / nb, "core": new}


# --------------------------------------------------------------------------- sweep driver


def accumulating_swipe(cores, x, y, loss="square", batch_size=-1, num_swipes=1, lr=1.0, method="exact",
                       eps=1e-12, eps_decay=None, orthonormalize=False, skip_second=False, direction="l2r",
                       loss_kwargs=None, trace=None, eps_per_node=False, adaptive_step=False, max_norm=None):
    """Site order, eps schedule and turn-around skip of network.py:409-608.
    ``cores`` is updated in place.  ``trace`` (a list) receives one dict per site
    update: NS, k, eps, loss.  Returns True, or False on a singular system."""
    n = len(cores)
    order = list(range(n))
    NS = 0
    last_l2r = None
    last_r2l = None

    def eps_at(NS_):
        e = eps[NS_] if isinstance(eps, list) else eps
        if eps_decay is not None:
            e = e * eps_decay ** NS_
        return e

    def one(k, e, left):
        try:
            r = site_update(cores, x, y, k, loss=loss, batch_size=batch_size, method=("exact" if (e == 0 and method == "ridge_exact") else method),
                            eps=e, lr=lr, loss_kwargs=loss_kwargs, adaptive_step=adaptive_step, max_norm=max_norm)
        except np.linalg.LinAlgError:
            return False
        if orthonormalize:
            (orthonormalize_left if left else orthonormalize_right)(cores, k)
        if trace is not None:
            trace.append({"NS": NS, "k": k, "eps": e, "loss": r["loss"], "A": r["A"], "b": r["b"], "step": r["step"]})
        return True

    for _ in range(num_swipes):
        e = eps_at(NS)
        first = order if direction == "l2r" else order[::-1]
        for i, k in enumerate(first):
            if eps_per_node and isinstance(eps, list):
                e = eps[i if direction == "l2r" else len(first) - 1 - i]
            if last_r2l is not None and k == last_r2l:
                last_l2r = k
                continue
            last_l2r = k
            if not one(k, e, True):
                return False
        NS += 1
        if skip_second:
            continue
        e = eps_at(NS)
        second = order[::-1] if direction == "l2r" else order
        for i, k in enumerate(second):
            if eps_per_node and isinstance(eps, list):
                e = eps[i if direction == "r2l" else len(second) - 1 - i]
            if last_l2r is not None and k == last_l2r:
                last_r2l = k
                continue
            last_r2l = k
            if not one(k, e, False):
                return False
        NS += 1
    return True


# --------------------------------------------------------------------------- CPD


def cpd_Z(factors, x):
    """Z_j[s,b(,o)] = sum_p x[s,p] A_j[b,p(,o)].  network.py:947-953."""
    xs = site_inputs(x, len(factors))
    return [np.tensordot(xs[j], factors[j], axes=(1, 1)) for j in range(len(factors))]


def cpd_forward(factors, x):
    """y[s,o] = sum_b prod_j Z_j[s,b].  network.py:961-974."""
    Z = cpd_Z(factors, x)
    prod = Z[0]  # (S, b, o)
    for z in Z[1:]:
        prod = prod * z[:, :, None]
    return prod.sum(axis=1)


def cpd_jacobian(factors, x, i):
    """J[s, o, (b,p[,o''])] for factor i.  network.py:955-959: for i>0 the einsum
    drops the output leg 'o' of Z_1 by summing it (quirk kept; exact for o=1)."""
    xs = site_inputs(x, len(factors))
    Z = cpd_Z(factors, x)
    S = xs[0].shape[0]
    O = factors[0].shape[2]
    Rk = factors[0].shape[0]
    f = factors[i].shape[1]
    if i == 0:
        other = np.ones((S, Rk))
        for j in range(1, len(factors)):
            other = other * Z[j]
        base = np.einsum("sb,sp->sbp", other, xs[0])
        J = np.zeros((S, O, Rk, f, O))
        for o in range(O):
            J[:, o, :, :, o] = base
        return J.reshape(S, O, Rk * f * O)
    other = Z[0].sum(axis=2)
    for j in range(1, len(factors)):
        if j != i:
            other = other * Z[j]
    base = np.einsum("sb,sp->sbp", other, xs[i]).reshape(S, 1, Rk * f)
    return np.broadcast_to(base, (S, O, Rk * f)).copy()


def cpd_site_update(factors, x, y, i, loss="square", batch_size=-1, method="ridge_cholesky", eps=0.0, lr=1.0,
                    loss_kwargs=None, apply=True):
    xs_all = site_inputs(x, len(factors))
    N = xs_all[0].shape[0]
    bs = N if batch_size <= 0 else batch_size
    nb = (N + bs - 1) // bs
    lf = LOSSES[loss] if isinstance(loss, str) else loss
    kw = loss_kwargs or {}
    A_out = b_out = None
    tot = 0.0
    y = np.asarray(y, dtype=np.float64)
    for bi in range(nb):
        sl = slice(bi * bs, (bi + 1) * bs)
        xb = [t[sl] for t in xs_all]
        pred = cpd_forward(factors, xb)
        lo, g, H = lf(pred, y[sl].reshape(pred.shape[0], -1), **kw)
        J = cpd_jacobian(factors, xb, i)
        A, b = gram(J, g, H)
        A_out = A if A_out is None else A_out + A
        b_out = b if b_out is None else b_out + b
        tot += float(np.mean(lo))
    theta = np.asarray(factors[i], dtype=np.float64)
    step = solve_system(A_out, b_out, theta.reshape(-1), method=method, eps=eps).reshape(theta.shape)
    new = update_node(theta, step, lr=lr)
    if apply:
        factors[i] = new
    return {"A": A_out, "b": b_out, "step": step, "loss": tot / nb, "core": new}


# --------------------------------------------------------------------------- matrix-free pieces


def matvec(J, H, v):
    """Av = J^T (H (J v)).  network.py:770-790."""
    S, C, P = J.shape
    coeff = np.einsum("scd,sdP,P->sc", H, J, v.reshape(P), optimize=True)
    return np.einsum("scP,sc->P", J, coeff, optimize=True)


def lanczos_solve(mv, b, x0, max_iter, tol):
    """Lanczos-Galerkin solve of A x = b from x0.  network.py:796-824."""
    v = [np.zeros_like(x0)]
    a = [0.0]
    bc = [0.0]
    r0 = b - mv(x0)
    beta1 = np.linalg.norm(r0)
    bc.append(beta1)
    v.append(r0 / beta1)
    j = 0
    for j in range(1, max_iter + 1):
        w = mv(v[j]) - bc[j] * v[j - 1]
        aj = float((w * v[j]).sum())
        a.append(aj)
        w = w - aj * v[j]
        bj = np.linalg.norm(w)
        bc.append(bj)
        v.append(w / bj)
        if bj < tol:
            break
    Vm = np.stack(v[1:j + 1], axis=-1)
    Tm = np.diag(np.array(a[1:]))
    if len(a) > 2:
        off = np.array(bc[2:j + 1])
        Tm = Tm + np.diag(off, 1) + np.diag(off, -1)
    rhs = np.zeros(len(a) - 1)
    rhs[0] = beta1
    yv = np.linalg.solve(Tm, rhs)
    return x0 + Vm @ yv
