"""CPU restatement (numpy) of the reference's patch/pixel conv-TT layer under its matrix-free sweeps.

TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product path.

Follows TensorConvolutionTrainLayer (reference tensor/layers.py:791-890: per column an input x[s,q,t], a pixel core
C_k[a,t,b], a patch core A_k[r,(c),q,r']; train-node order C1, A1, C2, A2, ...) contracted the way
TensorNetwork.compute_stacks / compute_jacobian_stack / get_J / get_b do (tensor/network.py:55-113, 219-291), and the Krylov
sweeps lanczos_swipe / scipy_swipe (tensor/network.py:709-932).  Everything is dense einsum on small cases.
Pinned against recordings of the reference itself: tests/golden/conv_*.npz (tests/golden/make_golden_conv.py).
"""
import numpy as np

from . import tn_oracle as orc


def canon_cores(cores, names, C):
    """Split the train-node list [C1, A1, C2, A2, ...] into canonical A_k (r, c, Q, r') and C_k (a, T, a') arrays."""
    n = len(cores) // 2
    A, Cc = [], []
    for k in range(n):
        c_t, a_t = cores[2 * k], cores[2 * k + 1]
        assert names[2 * k] == f"C{k + 1}" and names[2 * k + 1] == f"A{k + 1}"
        if k == 0:
            c3 = c_t.reshape(1, c_t.shape[0], -1)
            a4 = a_t.reshape(1, C, a_t.shape[-2], a_t.shape[-1]) if n > 1 else a_t.reshape(1, C, -1, 1)
        elif k == n - 1:
            c3 = c_t.reshape(c_t.shape[0], c_t.shape[1], 1)
            a4 = a_t.reshape(a_t.shape[0], 1, a_t.shape[1], 1)
        else:
            c3 = c_t
            a4 = a_t.reshape(a_t.shape[0], 1, a_t.shape[1], a_t.shape[2])
        A.append(a4)
        Cc.append(c3)
    return A, Cc


def column(A4, C3, x):
    """W[s, c, (a, r), (b, r')] = sum_{q,t} A[r,c,q,r'] C[a,t,b] x[s,q,t]: the per-sample transfer tensor of one column."""
    return np.einsum("rcqu,atb,sqt->scarbu", A4, C3, x, optimize=True)


def forward(A, Cc, x):
    """yhat[s, c] (tensor/network.py:115-137)."""
    S = x.shape[0]
    C = A[0].shape[1]
    E = np.ones((S, C, 1, 1))
    for A4, C3 in zip(A, Cc):
        W = column(A4, C3, x)
        if W.shape[1] == 1 and C > 1:
            W = np.broadcast_to(W, (S, C) + W.shape[2:])
        E = np.einsum("scar,scarbu->scbu", E, W, optimize=True)
    return E[:, :, 0, 0]


def envs(A, Cc, x):
    """Left environments E_k (s, C, a', r') after column k and right environments R_k (s, a, r) of columns k..n-1."""
    S = x.shape[0]
    C = A[0].shape[1]
    n = len(A)
    L, R = [None] * n, [None] * (n + 1)
    E = np.ones((S, C, 1, 1))
    for k in range(n):
        W = column(A[k], Cc[k], x)
        if W.shape[1] == 1 and C > 1:
            W = np.broadcast_to(W, (S, C) + W.shape[2:])
        E = np.einsum("scar,scarbu->scbu", E, W, optimize=True)
        L[k] = E
    Rt = np.ones((S, 1, 1))
    R[n] = Rt
    for k in range(n - 1, 0, -1):
        W = column(A[k], Cc[k], x)[:, 0]
        Rt = np.einsum("sarbu,sbu->sar", W, Rt, optimize=True)
        R[k] = Rt
    return L, R


def jacobian(A, Cc, x, kind, k):
    """J[s, c, P] of the prediction w.r.t. node (kind, k), P in the node's own row-major order (network.py:101-113)."""
    S = x.shape[0]
    C = A[0].shape[1]
    n = len(A)
    L, R = envs(A, Cc, x)
    El = L[k - 1] if k > 0 else np.ones((S, C, 1, 1))
    Rr = R[k + 1]
    if kind == "A":
        Y = np.einsum("sqt,atb->sqab", x, Cc[k], optimize=True)
        if k == 0:
            base = np.einsum("sqab,sbu->squ", Y, Rr, optimize=True)           # a == 1
            J = np.zeros((S, C, C) + base.shape[1:])
            for c in range(C):
                J[:, c, c] = base
            return J.reshape(S, C, -1)
        J = np.einsum("scar,sqab,sbu->scrqu", El, Y, Rr, optimize=True)
        return J.reshape(S, C, -1)
    if k == 0:
        J = np.einsum("cqu,sqt,sbu->sctb", A[0][0], x, Rr, optimize=True)      # a == 1: P = (t, b)
        return J.reshape(S, C, -1)
    J = np.einsum("scar,rqu,sqt,sbu->scatb", El, A[k][:, 0], x, Rr, optimize=True)
    return J.reshape(S, C, -1)


def node_list(n):
    return [(kind, k) for k in range(n) for kind in ("C", "A")]


def site_problem(cores, names, C, x, y, loss, idx, batch_size):
    """(mean-of-batch-means loss, b, [(J, H) per batch]) for train node idx, batched as lanczos_swipe does (network.py:740-765)."""
    A, Cc = canon_cores(cores, names, C)
    kind, k = node_list(len(A))[idx]
    N = x.shape[0]
    bs = N if batch_size <= 0 else batch_size
    b = None
    parts = []
    tot = 0.0
    nb = (N + bs - 1) // bs
    for i in range(nb):
        xb, yb = x[i * bs:(i + 1) * bs], y[i * bs:(i + 1) * bs]
        pred = forward(A, Cc, xb)
        lo, g, H = orc.LOSSES[loss](pred, yb)
        J = jacobian(A, Cc, xb, kind, k)
        bb = np.einsum("scP,sc->P", J, g)
        b = bb if b is None else b + bb
        parts.append((J, H))
        tot += float(np.mean(lo))
    return tot / nb, b, parts


def matvec_of(parts):
    return lambda v: sum(orc.matvec(J, H, v) for J, H in parts)


def lanczos_swipe(cores, names, C, x, y, loss, batch_size, num_swipes, lr, max_iter, tol, x0s, trace=None):
    """lanczos_swipe (network.py:709-832) with the recorded start vectors; returns the per-node losses."""
    cores = [c.copy() for c in cores]
    losses = []
    it = iter(x0s)
    for NS in range(num_swipes):
        order = list(range(len(cores))) if NS % 2 == 0 else list(reversed(range(len(cores))))
        for idx in order:
            lo, b, parts = site_problem(cores, names, C, x, y, loss, idx, batch_size)
            losses.append(lo)
            x0 = next(it).reshape(-1)
            step = orc.lanczos_solve(matvec_of(parts), -b, x0, max_iter, tol)
            cores[idx] = cores[idx] + lr * step.reshape(cores[idx].shape)
            if trace is not None:
                trace.append({"NS": NS, "k": idx, "after": [c.copy() for c in cores]})
    return cores, losses
