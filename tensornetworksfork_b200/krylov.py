"""On-device Krylov solvers for the matrix-free local solve (CG, MINRES) and a SciPy bridge.

``scipy_swipe`` of the reference hands a ``LinearOperator`` to ``scipy.sparse.linalg.cg`` / ``minres`` and
crosses the torch<->numpy boundary in float32 every iteration (tensor/network.py:896-926).  Here the
recurrences run in float64 on the device around ``ops.matvec``; the only host round trip per
iteration is the scalar convergence test.  ``scipy_bridge`` keeps the reference's exact behaviour
(float32 host recurrences) for callers that pass a SciPy solver object.
"""
import numpy as np
import torch


def cg(matvec, b, x0=None, maxiter=50, rtol=1e-6):
    """Conjugate gradients on A x = b (A symmetric positive semi-definite). Stops at ||r|| <= rtol*||b||."""
    x = torch.zeros_like(b) if x0 is None else x0.clone().reshape(-1)
    r = b - matvec(x) if x0 is not None else b.clone()
    p = r.clone()
    rs = torch.dot(r, r)
    bnorm = float(torch.norm(b).item())
    if bnorm == 0.0:
        return x
    for _ in range(maxiter):
        if float(rs.sqrt().item()) <= rtol * bnorm:
            break
        Ap = matvec(p)
        alpha = rs / torch.dot(p, Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rs_new = torch.dot(r, r)
        p = r + (rs_new / rs) * p
        rs = rs_new
    return x


def minres(matvec, b, x0=None, maxiter=50, rtol=1e-6):
    """MINRES (Paige & Saunders) for symmetric A; minimises ||b - A x|| over the Krylov space."""
    x = torch.zeros_like(b) if x0 is None else x0.clone().reshape(-1)
    r1 = b - matvec(x) if x0 is not None else b.clone()
    beta1 = float(torch.norm(r1).item())
    if beta1 == 0.0:
        return x
    bnorm = float(torch.norm(b).item())
    y = r1
    r2 = r1.clone()
    oldb, beta, dbar, epsln, phibar = 0.0, beta1, 0.0, 0.0, beta1
    cs, sn = -1.0, 0.0
    w = torch.zeros_like(b)
    w2 = torch.zeros_like(b)
    for itn in range(1, maxiter + 1):
        v = y / beta
        y = matvec(v)
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = float(torch.dot(v, y).item())
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        oldb = beta
        beta = float(torch.norm(r2).item())
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
        if phibar <= rtol * bnorm or beta == 0.0:
            break
    return x


def scipy_bridge(solver, matvec, b, x0=None, maxiter=50, rtol=1e-6):
    """Run a ``scipy.sparse.linalg`` solver exactly as the reference does: float32 vectors on the host,
    the matvec on the device (tensor/network.py:896-926)."""
    from scipy.sparse.linalg import LinearOperator
    dev, dt = b.device, b.dtype

    def mv(v):
        t = torch.tensor(v, dtype=dt, device=dev).reshape(-1)
        return matvec(t).flatten().float().cpu().numpy()

    b_np = b.flatten().float().cpu().numpy()
    op = LinearOperator((b_np.shape[0], b_np.shape[0]), matvec=mv)
    x_sol, _ = solver(op, b_np, x0=x0, maxiter=maxiter, rtol=rtol)
    return x_sol
