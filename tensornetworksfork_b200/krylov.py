"""Krylov solvers of the matrix-free local solve (CG, MINRES, Lanczos-Galerkin) and a SciPy bridge.

``scipy_swipe`` of the reference hands a ``LinearOperator`` to ``scipy.sparse.linalg.cg`` / ``minres`` and crosses the
torch<->numpy boundary in float32 every iteration (tensor/network.py:896-926); ``lanczos_swipe`` runs an eager torch loop
(:793-824).  Here the recurrences run in float64 inside libtn_b200.so (csrc/krylov.cu: ``tn_cg`` / ``tn_minres`` /
``tn_lanczos``): all scalars stay on the device and the host only polls a convergence flag.  ``matvec`` is either an
``ops.Operator`` (Kronecker factors: the built-in two-pass kernels) or any callable ``v -> A v`` (conv-TT, cum-sum), which the
drivers call back between their own kernels.  ``scipy_bridge`` keeps the reference's exact behaviour (float32 host recurrences)
for callers that pass a SciPy solver object.
"""
import torch

from . import ops


def as_operator(matvec, b):
    if isinstance(matvec, ops.Operator):
        return matvec
    return ops.Operator(b.numel(), matvec=matvec, device=b.device)


def cg(matvec, b, x0=None, maxiter=50, rtol=1e-6):
    """Conjugate gradients on A x = b (A symmetric positive semi-definite). Stops at ||r|| <= rtol*||b||."""
    return ops.cg(as_operator(matvec, b), b.contiguous().view(-1), x0=x0, max_iter=maxiter, rtol=rtol)[0]


def minres(matvec, b, x0=None, maxiter=50, rtol=1e-6):
    """MINRES (Paige & Saunders) for symmetric A; minimises ||b - A x|| over the Krylov space."""
    return ops.minres(as_operator(matvec, b), b.contiguous().view(-1), x0=x0, max_iter=maxiter, rtol=rtol)[0]


def lanczos(matvec, b, x0, maxiter=50, tol=1e-6):
    """Lanczos-Galerkin solve of A x = b started at x0 (reference tensor/network.py:793-824)."""
    return ops.lanczos(as_operator(matvec, b), b.contiguous().view(-1), x0=x0, max_iter=maxiter, tol=tol)[0]


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
