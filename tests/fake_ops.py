"""CPU stand-ins for tensornetworksfork_b200.ops, for exercising the HOST logic (sweep driver,
caching, sharding, all-reduce bookkeeping) without a GPU.  TEST INFRASTRUCTURE ONLY: each function
restates what the corresponding C-ABI kernel computes, with torch CPU ops, following the oracle."""
import numpy as np
import torch

from tensornetworksfork_b200 import ops as real
from tensornetworksfork_b200.ops import Factor, npairs  # noqa: F401


def _unit_rows(t, what):
    """The kernels read ``t`` as rows with stride ``t.stride(0)`` and unit stride inside a row (ops.Factor.ld, env_ld, dot_ld, out_ld)."""
    assert t is None or t.dim() == 1 or (t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1)), (what, tuple(t.shape), t.stride())


def _bare(*ts):
    """Buffers the kernels take as bare pointers (weights, outputs, vectors) must be dense."""
    for t in ts:
        assert t is None or t.is_contiguous(), (tuple(t.shape), t.stride())


def _rows(f: Factor, rows):
    t = f.tensor
    _unit_rows(t, "factor")
    assert t.dtype == torch.float64, t.dtype                     # ops._need_cuda: the kernels read fp64
    idx = torch.arange(rows) // max(f.div, 1)
    # the kernels read row (rows - 1) // div of the factor: it must exist (no clamping on the device)
    assert rows == 0 or (rows - 1) // max(f.div, 1) < t.shape[0], (rows, f.div, tuple(t.shape))
    need = (f.col + f.m) if f.map_kind == real.MAP_IDENTITY else (f.col + 1)
    assert t.dim() != 2 or need <= t.shape[1], (f.col, f.m, tuple(t.shape))
    if f.map_kind == real.MAP_IDENTITY:
        return t[idx][:, f.col:f.col + f.m]
    x = t[idx][:, f.col]
    if f.map_kind == real.MAP_SINCOS:
        return torch.stack([torch.cos(0.5 * np.pi * x), torch.sin(0.5 * np.pi * x)], dim=1)
    return torch.stack([x ** d for d in range(f.m)], dim=1)


def ones_factor(like):
    return Factor(torch.ones(1, 1, dtype=torch.float64), m=1, div=1 << 30)


def env_update(env_in, x, core3, rows, cdiv=1, env_div=1, out=None):
    _unit_rows(env_in, "env_in")
    _unit_rows(out, "out")
    assert env_in is None or (env_in.dtype == torch.float64 and (rows - 1) // env_div < env_in.shape[0] and env_in.shape[1] >= core3.shape[0])
    assert core3.dtype == torch.float64 and core3.dim() == 3
    assert env_in is not None or core3.shape[0] == 1            # tn_env_update: env_in == NULL requires r_in == 1
    assert x.map_kind != real.MAP_SINCOS or core3.shape[1] == 2  # sin-cos map has f == 2
    phi = _rows(Factor(x.tensor, m=x.m, div=cdiv, map_kind=x.map_kind, col=x.col), rows)
    e = torch.ones(rows, 1, dtype=torch.float64) if env_in is None else env_in[torch.arange(rows) // env_div]
    res = torch.einsum("sa,sp,apb->sb", e, phi, core3)
    if out is not None:
        out.copy_(res)
        return out
    return res


def predict(env_in, x, core3, dot, rows, cdiv=1, env_div=1, dot_div=1, out=None):
    _unit_rows(dot, "dot")
    assert dot.dtype == torch.float64 and dot.shape[-1] >= core3.shape[2] and (dot_div >= (1 << 30) or (rows - 1) // dot_div < dot.shape[0])
    o = env_update(env_in, x, core3, rows, cdiv, env_div, out=None)
    d = dot[(torch.arange(rows) // dot_div).clamp(max=dot.shape[0] - 1)]
    y = (o * d).sum(1)
    if out is not None:
        out.copy_(y)
        return out
    return y


def class_rows(env, U, g):
    F = torch.einsum("stc,sca->sta", U, env).reshape(-1, env.shape[2]) if U is not None else None
    G = torch.einsum("sc,sca->sa", g, env) if g is not None else None
    return F, G


def _pairs(F):
    m = F.shape[1]
    return torch.stack([F[:, i] * F[:, j] for i in range(m) for j in range(i, m)], dim=1)


def gram(mode, fa, fb, fc, w, rows, M=None, accumulate=False, flush_rows=None):
    _bare(w, M)
    A, B, C = _pairs(_rows(fa, rows)), _pairs(_rows(fb, rows)), _pairs(_rows(fc, rows))
    ww = torch.ones(rows, dtype=torch.float64) if w is None else w
    out = torch.einsum("s,sa,sb,sc->abc", ww, A, B, C).reshape(-1)
    if M is None:
        return out
    if accumulate:
        M += out
    else:
        M.copy_(out)
    return M


def rhs(fa, fb, fc, w, rows, b=None, accumulate=False):
    _bare(w, b)
    ww = torch.ones(rows, dtype=torch.float64) if w is None else w
    out = torch.einsum("s,sa,sb,sc->abc", ww, _rows(fa, rows), _rows(fb, rows), _rows(fc, rows)).reshape(-1)
    if b is None:
        return out
    if accumulate:
        b += out
    else:
        b.copy_(out)
    return b


def gram_generic(f1, f2, f3, t1, t2, t3, w, rows, rhs_only=False, out=None, accumulate=False):
    _bare(w, out, t1, t2, t3)
    J = _rows(f1, rows)[:, t1.long()] * _rows(f2, rows)[:, t2.long()] * _rows(f3, rows)[:, t3.long()]
    ww = torch.ones(rows, dtype=torch.float64) if w is None else w
    res = (J.t() @ ww) if rhs_only else ((J * ww[:, None]).t() @ J).reshape(-1)
    if out is None:
        return res
    if accumulate:
        out += res
    else:
        out.copy_(res)
    return out


def _dense(M, m_pos, role_of_pos):
    n = [npairs(m) for m in m_pos]
    n_role = [0, 0, 0]
    for t in range(3):
        n_role[role_of_pos[t]] = n[t]
    Mt = M.reshape(n_role)
    P = m_pos[0] * m_pos[1] * m_pos[2]
    idx = torch.arange(P)
    ii = [idx // (m_pos[1] * m_pos[2]), (idx // m_pos[2]) % m_pos[1], idx % m_pos[2]]
    q = [None, None, None]
    for t in range(3):
        lo = torch.minimum(ii[t][:, None], ii[t][None, :])
        hi = torch.maximum(ii[t][:, None], ii[t][None, :])
        q[role_of_pos[t]] = lo * m_pos[t] - (lo * (lo - 1)) // 2 + (hi - lo)
    return Mt[q[0], q[1], q[2]]


def gram_sigma(M, m_pos, role_of_pos):
    s = _dense(M, m_pos, role_of_pos).diagonal().abs().mean()
    return torch.where(s == 0, torch.ones_like(s), s).reshape(1)


def gram_expand(M, m_pos, role_of_pos, sigma, ridge, A=None):
    D = _dense(M, m_pos, role_of_pos) / sigma
    D = D + ridge * torch.eye(D.shape[0], dtype=torch.float64)
    return D.contiguous()


def rhs_prepare(b, theta, sigma, ridge):
    v = b / sigma
    if ridge != 0:
        v = v + ridge * theta
    return -v


def cholesky_solve(A, rhs_vec):
    P = A.shape[0]
    info = torch.zeros(1, dtype=torch.int32)
    try:
        L = torch.linalg.cholesky(A[:, :P])
    except torch.linalg.LinAlgError:
        info[0] = 1
        return info
    rhs_vec.copy_(torch.cholesky_solve(rhs_vec.unsqueeze(-1), L).squeeze(-1))
    return info


def cholesky_solve_mixed(A, rhs_vec, rtol=1e-12, max_iter=12):
    info = cholesky_solve(A, rhs_vec)
    return info, torch.tensor([0.0, 0.0], dtype=torch.float64)


def update_node(theta, step, lr=1.0, adaptive_step=False, max_norm=None):
    if adaptive_step:
        sn, pn = torch.norm(step), torch.norm(theta)
        if sn > pn:
            step = step * (pn / sn)
    theta += lr * step
    if max_norm is not None:
        cn = torch.norm(theta)
        if cn > max_norm:
            theta *= max_norm / cn
    return theta


def qr(a):
    assert a.is_contiguous() and a.dim() == 2
    Q, R = torch.linalg.qr(a, mode="reduced")
    a.copy_(Q)
    return R.contiguous()        # the real entry point returns a dense row-major R


def matvec(fa, fb, fc, w, rows, v, out=None):
    _bare(w, out)
    J = torch.einsum("sa,sb,sc->sabc", _rows(fa, rows), _rows(fb, rows), _rows(fc, rows)).reshape(rows, -1)
    ww = torch.ones(rows, dtype=torch.float64) if w is None else w
    return J.t() @ (ww * (J @ v))


def cholesky_factor(A, tensor_core=False):
    P = A.shape[0]
    info = torch.zeros(1, dtype=torch.int32)
    try:
        L = torch.linalg.cholesky(A[:, :P])
    except torch.linalg.LinAlgError:
        info[0] = 1
        return torch.zeros(1, dtype=torch.float64), info
    A[:, :P] = torch.tril(L) + torch.triu(A[:, :P], 1)       # the factor overwrites the lower triangle only
    return torch.zeros(1, dtype=torch.float64), info


def cholesky_apply(L, work, info, x):
    P = L.shape[0]
    if int(info[0]) == 0:
        x.copy_(torch.cholesky_solve(x.unsqueeze(-1), torch.tril(L[:, :P])).squeeze(-1))
    return x


def gram_trace(fa, fb, fc, w, rows, out=None):
    _bare(w, out)
    q = (_rows(fa, rows) ** 2).sum(1) * (_rows(fb, rows) ** 2).sum(1) * (_rows(fc, rows) ** 2).sum(1)
    ww = torch.ones(rows, dtype=torch.float64) if w is None else w
    res = torch.stack([(ww * q).sum(), (ww.abs() * q).sum()])
    if out is not None:
        out.copy_(res)
        return out
    return res


def _apply(op, v):
    out = op(v)
    if op.sigma is not None:
        out = out / op.sigma
    return out + op.ridge * v if op.ridge != 0.0 else out


def cg(op, b, x0=None, precond=None, max_iter=50, rtol=1e-6, poll_every=None):
    """Stand-in of tn_cg: the recurrences of csrc/krylov.cu in torch."""
    L, lwork, linfo = precond if precond is not None else (None, None, None)
    stats = torch.zeros(4, dtype=torch.float64)
    if linfo is not None and int(linfo[0]) != 0:
        stats[2] = 1.0
        return (torch.zeros_like(b) if x0 is None else x0.clone().reshape(-1)), stats

    def prec(r):
        return r if L is None else cholesky_apply(L, lwork, linfo, r.clone())

    if x0 is not None:
        x = x0.clone().reshape(-1)
    elif L is not None:
        x = prec(b)
    else:
        x = torch.zeros_like(b)
    r = b - _apply(op, x) if (x0 is not None or L is not None) else b.clone()
    bn = float(torch.dot(b, b))
    stats = torch.zeros(5, dtype=torch.float64)
    rz_old, p, it = None, None, 0
    for it in range(max_iter + 1):
        rel = float(torch.sqrt(torch.dot(r, r) / bn)) if bn > 0 else 0.0
        z = prec(r)
        crit = rel
        if L is not None:
            xx, zz = float(torch.dot(x, x)), float(torch.dot(z, z))
            crit = (zz / xx) ** 0.5 if xx > 0 else (1.0 if zz > 0 else 0.0)
            if not bn > 0:
                crit = 0.0
        stats[0], stats[1], stats[4] = rel, it, crit
        if not crit > rtol:
            stats[2] = 1.0
            break
        if it == max_iter:
            break
        rz = torch.dot(r, z)
        p = z if p is None else z + (rz / rz_old) * p
        q = _apply(op, p)
        alpha = rz / torch.dot(p, q)
        x = x + alpha * p
        r = r - alpha * q
        rz_old = rz
    stats[3] = op.applies
    return x, stats


def minres(op, b, x0=None, max_iter=50, rtol=1e-6, poll_every=None):
    """Stand-in of tn_minres (Paige & Saunders, as scipy.sparse.linalg.minres without preconditioner or shift)."""
    stats = torch.zeros(4, dtype=torch.float64)
    x = torch.zeros_like(b) if x0 is None else x0.clone().reshape(-1)
    r1 = b - _apply(op, x) if x0 is not None else b.clone()
    beta1 = float(torch.norm(r1))
    if beta1 == 0.0:
        stats[2] = 1.0
        return x, stats
    bnorm = float(torch.norm(b))
    y = r1
    r2 = r1.clone()
    oldb, beta, dbar, epsln, phibar = 0.0, beta1, 0.0, 0.0, beta1
    cs, sn = -1.0, 0.0
    w = torch.zeros_like(b)
    w2 = torch.zeros_like(b)
    for itn in range(1, max_iter + 1):
        v = y / beta
        y = _apply(op, v)
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = float(torch.dot(v, y))
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        oldb = beta
        beta = float(torch.norm(r2))
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        gamma = max(float(np.hypot(gbar, beta)), 1e-300)
        cs, sn = gbar / gamma, beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) / gamma
        x = x + phi * w
        stats[0], stats[1] = (phibar / bnorm if bnorm > 0 else 0.0), itn
        if phibar <= rtol * bnorm or beta == 0.0:
            stats[2] = 1.0
            break
    stats[3] = op.applies
    return x, stats


def lanczos(op, b, x0=None, max_iter=50, tol=1e-6, poll_every=None):
    """Stand-in of tn_lanczos: the reference's own recurrence (tensor/network.py:793-824)."""
    stats = torch.zeros(4, dtype=torch.float64)
    x0v = torch.zeros_like(b) if x0 is None else x0.reshape(-1)
    vs = [torch.zeros_like(b)]
    alphas, betas = [], [None]
    r0 = b - _apply(op, x0v) if x0 is not None else b.clone()
    beta1 = torch.norm(r0)
    betas.append(beta1)
    vs.append(r0 / beta1)
    j = 0
    for j in range(1, max_iter + 1):
        wv = _apply(op, vs[j]) - betas[j] * vs[j - 1] if j > 1 else _apply(op, vs[j])
        a_j = (wv * vs[j]).sum()
        alphas.append(a_j)
        wv = wv - a_j * vs[j]
        b_j = torch.norm(wv)
        betas.append(b_j)
        vs.append(wv / b_j)
        if float(b_j) < tol:
            stats[2] = 1.0
            break
    Vm = torch.stack(vs[1:j + 1], dim=-1)
    Tm = torch.diag(torch.stack(alphas))
    if len(alphas) > 1:
        off = torch.stack(betas[2:j + 1])
        Tm = Tm + torch.diag(off, 1) + torch.diag(off, -1)
    e1 = torch.zeros(len(alphas), dtype=b.dtype)
    e1[0] = beta1
    yv = torch.linalg.solve(Tm, e1)
    stats[0], stats[1], stats[3] = float(betas[-1]), j, op.applies
    return x0v + Vm @ yv, stats


NAMES = ["ones_factor", "env_update", "predict", "class_rows", "gram", "rhs", "gram_generic", "gram_sigma", "gram_expand", "rhs_prepare",
         "cholesky_solve", "cholesky_solve_mixed", "update_node", "qr", "matvec", "bmm", "outer_rows", "rows_dot", "cholesky_factor",
         "cholesky_apply", "gram_trace", "cg", "minres", "lanczos"]


def install(monkeypatch=None):
    """Swap the kernels for the stand-ins and lift the CUDA-only guard (tests only)."""
    import sys
    from tensornetworksfork_b200.tensor import network
    me = sys.modules[__name__]
    for n in NAMES:
        if monkeypatch is not None:
            monkeypatch.setattr(real, n, getattr(me, n))
        else:
            setattr(real, n, getattr(me, n))
    from tensornetworksfork_b200.tensor import conv
    for cls in (network.TensorNetwork, conv.ConvTrainNetwork):
        if monkeypatch is not None:
            monkeypatch.setattr(cls, "_require_cuda", lambda self, dev: None)
        else:
            cls._require_cuda = lambda self, dev: None


def bmm(A, B, out=None, accumulate=False):
    assert out is None or out.is_contiguous()
    assert 8 * (A.shape[-2] * A.shape[-1] + B.shape[-2] * B.shape[-1]) <= 200 * 1024   # tn_bmm: one sample's operands fit shared memory
    r = torch.matmul(A, B)
    if r.dim() == 2:
        r = r.unsqueeze(0)
    if out is None:
        return r.contiguous()
    if accumulate:
        out += r
    else:
        out.copy_(r)
    return out


def outer_rows(G, W, w=None, gdiv=1, out=None, accumulate=False):
    # the layout contract of the real entry point (ops.outer_rows): the stand-in must refuse what the kernel would misread
    assert G.dim() == 2 and W.dim() == 2 and G.stride(1) == 1 and W.stride(1) == 1, (G.stride(), W.stride())
    _bare(w)
    assert out is None or out.is_contiguous()
    assert 1 <= G.shape[1] <= 128, G.shape                      # tn_outer_rows: ra <= OR_MAXRA
    rows = W.shape[0]
    Gr = G[torch.arange(rows) // gdiv]
    if w is not None:
        Gr = Gr * w[:, None]
    r = Gr.t() @ W
    if out is None:
        return r.contiguous()
    if accumulate:
        out += r
    else:
        out.copy_(r)
    return out


def rows_dot(W, V, out=None):
    assert W.dim() == 2 and V.dim() == 2 and W.stride(1) == 1 and V.stride(1) == 1 and W.shape[1] == V.shape[1], (W.stride(), V.stride())
    assert out is None or (out.dim() == 2 and (out.stride(1) == 1 or V.shape[0] == 1))
    r = W @ V.t()
    if out is None:
        return r.contiguous()
    out.copy_(r)
    return out
