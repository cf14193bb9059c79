"""Torch-tensor front end of the C ABI: borrows device pointers for the duration of a call.

PyTorch is used here for device memory and streams only; every function below enqueues the
hand-written kernels of libtn_b200.so on torch's current CUDA stream.  No function has a CPU
implementation: a non-CUDA tensor raises.
"""
import ctypes
import functools
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import tn_factor

MAP_IDENTITY, MAP_SINCOS, MAP_POLY = 0, 1, 2
GRAM_FP64, GRAM_TF32, GRAM_TF32X3, GRAM_F16 = 0, 1, 2, 3


@dataclass
class Factor:
    """One Kronecker factor of the local Jacobian: rows of ``tensor`` (2-D, last dim contiguous)."""
    tensor: torch.Tensor
    m: int
    div: int = 1
    map_kind: int = MAP_IDENTITY
    col: int = 0  # first column read (raw feature index for SINCOS / POLY)

    def __post_init__(self):
        _unit_rows(self.tensor, "factor")

    @property
    def ld(self):
        return self.tensor.stride(0) if self.tensor.dim() == 2 else 1

    def ptr(self):
        return self.tensor.data_ptr() + 8 * self.col

    def c(self):
        return tn_factor(self.ptr(), self.ld, self.m, self.div, self.map_kind, 0)


def _unit_rows(t, what="operand"):
    """Layout contract of every row-indexed operand (factors, environments, outputs): at most 2-D, unit stride inside a row, a
    non-negative row stride (0 = one row shared by all samples, which the kernels only ever pair with a huge divisor) -- a transposed
    or broadcast view would be read out of layout, so it is refused here instead."""
    if t is None:
        return
    if t.dim() > 2 or (t.dim() >= 1 and t.shape[-1] > 1 and t.stride(-1) != 1) or (t.dim() == 2 and t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        raise _lib.TnError(f"{what}: rows with unit inner stride expected, got shape {tuple(t.shape)} with strides {t.stride()}")


def _need_cuda(*ts):
    """All operands are fp64 CUDA tensors on ONE device; returns that device."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.TnError("tensornetworksfork_b200 kernels need CUDA tensors; there is no CPU path")
        if t.dtype != torch.float64:
            raise _lib.TnError(f"fp64 tensors expected, got {t.dtype}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _lib.TnError(f"operands live on different devices ({dev} and {t.device})")
    return dev


class _on:
    """Make `dev` the current CUDA device for the duration of a call (kernels launch on the calling thread's current device, and
    the stream handed to the library has to belong to it); free when it already is."""

    def __init__(self, dev):
        self.dev = dev
        self.ctx = None

    def __enter__(self):
        if self.dev is not None and self.dev.index is not None and self.dev.index != torch.cuda.current_device():
            self.ctx = torch.cuda.device(self.dev)
            self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        return False


def _op(fn):
    """Run a wrapper with the operands' device current (a process driving several GPUs launched on the wrong one)."""

    @functools.wraps(fn)
    def inner(*a, **k):
        dev = None
        for v in list(a) + list(k.values()):
            t = v.tensor if isinstance(v, Factor) else v
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if dev is None:
                    dev = t.device
                elif t.device != dev:
                    raise _lib.TnError(f"operands live on different devices ({dev} and {t.device})")
        with _on(dev):
            return fn(*a, **k)

    return inner


def _bare(*ts):
    """Buffers handed to the library as bare pointers (weights, vectors, outputs) must be dense: the kernels index them linearly."""
    for t in ts:
        if t is not None and not t.is_contiguous():
            raise _lib.TnError(f"dense tensor expected, got shape {tuple(t.shape)} with strides {t.stride()}")


def _system_layout(A, rhs_vec, P):
    if A.dim() != 2 or A.stride(1) != 1 or A.stride(0) < P or A.shape[1] < P:
        raise _lib.TnError(f"system matrix must be ({P}, >= {P}) with unit inner stride, got {tuple(A.shape)} / {A.stride()}")
    if rhs_vec.numel() != P or not rhs_vec.is_contiguous():
        raise _lib.TnError(f"right-hand side must be a dense vector of {P}, got {tuple(rhs_vec.shape)} / {rhs_vec.stride()}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def ones_factor(like):
    return Factor(torch.ones(1, 1, dtype=torch.float64, device=like.device), m=1, div=1 << 30)


@_op
def env_update(env_in, x: Factor, core3, rows, cdiv=1, env_div=1, out=None):
    """out[row,b] = sum_{a,p} env_in[row/env_div,a] phi[row/cdiv,p] core3[a,p,b]."""
    lib = _lib.load()
    r_in, f, r_out = core3.shape
    core3 = core3.contiguous()
    _need_cuda(env_in, x.tensor, core3, out)
    _unit_rows(env_in, "env_in")
    _unit_rows(out, "out")
    if out is None:
        out = torch.empty((rows, r_out), dtype=torch.float64, device=core3.device)
    env_ld = env_in.stride(0) if env_in is not None else 0
    rc = lib.tn_env_update(_p(env_in), env_ld, env_div, ctypes.c_void_p(x.ptr()), x.ld, x.map_kind, f, cdiv, _p(core3),
                           _p(out), out.stride(0), None, 0, 1, None, rows, r_in, r_out, _stream())
    _lib.check(rc, "tn_env_update")
    return out


@_op
def predict(env_in, x: Factor, core3, dot, rows, cdiv=1, env_div=1, dot_div=1, out=None):
    """yhat[row] = sum_{a,p,b} env_in[row/env_div,a] phi[row/cdiv,p] core3[a,p,b] dot[row/dot_div,b]."""
    lib = _lib.load()
    r_in, f, r_out = core3.shape
    core3 = core3.contiguous()
    _need_cuda(env_in, x.tensor, core3, dot, out)
    _unit_rows(env_in, "env_in")
    _unit_rows(dot, "dot")
    _bare(out)
    if out is None:
        out = torch.empty((rows,), dtype=torch.float64, device=core3.device)
    env_ld = env_in.stride(0) if env_in is not None else 0
    rc = lib.tn_env_update(_p(env_in), env_ld, env_div, ctypes.c_void_p(x.ptr()), x.ld, x.map_kind, f, cdiv, _p(core3),
                           None, 0, _p(dot), dot.stride(0), dot_div, _p(out), rows, r_in, r_out, _stream())
    _lib.check(rc, "tn_env_update(predict)")
    return out


@_op
def class_rows(env, U, g):
    """F[(s,t),a] = sum_c U[s,t,c] env[s,c,a];  G[s,a] = sum_c g[s,c] env[s,c,a]."""
    lib = _lib.load()
    S, C, r = env.shape
    env = env.contiguous()
    _need_cuda(env, U, g)
    F = G = None
    V = 1
    if U is not None:
        U = U.contiguous()
        V = U.shape[1]
        F = torch.empty((S * V, r), dtype=torch.float64, device=env.device)
    if g is not None:
        g = g.contiguous()
        G = torch.empty((S, r), dtype=torch.float64, device=env.device)
    rc = lib.tn_class_rows(_p(env), _p(U), _p(g), _p(F), _p(G), S, C, V, r, _stream())
    _lib.check(rc, "tn_class_rows")
    return F, G


def npairs(m):
    return m * (m + 1) // 2


@_op
def gram(mode, fa: Factor, fb: Factor, fc: Factor, w, rows, M=None, accumulate=False, flush_rows=None):
    """M[qa,qb,qc] (+)= sum_rows w * pair(fa)[qa] * pair(fb)[qb] * pair(fc)[qc].  ``flush_rows``: fp32 accumulation window of the
    tensor-core modes for this call (None = the library default)."""
    lib = _lib.load()
    if flush_rows is not None and mode != GRAM_FP64:
        prev = lib.tn_gram_tc_flush_rows(int(flush_rows))
        try:
            return _gram(lib, mode, fa, fb, fc, w, rows, M, accumulate)
        finally:
            lib.tn_gram_tc_flush_rows(prev)
    return _gram(lib, mode, fa, fb, fc, w, rows, M, accumulate)


def _gram(lib, mode, fa, fb, fc, w, rows, M, accumulate):
    _need_cuda(fa.tensor, fb.tensor, fc.tensor, w)
    _bare(w, M)
    n = npairs(fa.m) * npairs(fb.m) * npairs(fc.m)
    dev = fa.tensor.device
    if M is None:
        M = torch.empty((n,), dtype=torch.float64, device=dev)
        accumulate = False
    ks = lib.tn_gram_ksplit(rows, fa.m, fb.m, fc.m, mode)
    work = None
    if mode == GRAM_FP64 and (ks > 1 or accumulate):
        work = torch.empty((ks * n,), dtype=torch.float64, device=dev)
    a, b, c = fa.c(), fb.c(), fc.c()
    rc = lib.tn_gram_kr3(mode, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), _p(w), rows, _p(M), _p(work), ks,
                         1 if accumulate else 0, _stream())
    _lib.check(rc, "tn_gram_kr3")
    return M


@_op
def rhs(fa: Factor, fb: Factor, fc: Factor, w, rows, b=None, accumulate=False):
    """b[ia,ib,ic] (+)= sum_rows w * fa[ia] fb[ib] fc[ic]."""
    lib = _lib.load()
    _need_cuda(fa.tensor, fb.tensor, fc.tensor, w)
    _bare(w, b)
    n = fa.m * fb.m * fc.m
    dev = fa.tensor.device
    if b is None:
        b = torch.empty((n,), dtype=torch.float64, device=dev)
        accumulate = False
    ks = lib.tn_rhs_ksplit(rows, fa.m, fb.m, fc.m)
    work = torch.empty((ks * n,), dtype=torch.float64, device=dev) if (ks > 1 or accumulate or n <= 4096) else None
    a, bb, c = fa.c(), fb.c(), fc.c()
    rc = lib.tn_rhs_kr3(ctypes.byref(a), ctypes.byref(bb), ctypes.byref(c), _p(w), rows, _p(b), _p(work), ks,
                        1 if accumulate else 0, _stream())
    _lib.check(rc, "tn_rhs_kr3")
    return b


@_op
def gram_generic(f1: Factor, f2: Factor, f3: Factor, t1, t2, t3, w, rows, rhs_only=False, out=None, accumulate=False):
    """Dense Gram J^T diag(w) J (P x P) or right-hand side J^T w (P) of J[:, i] = f1[t1[i]] f2[t2[i]] f3[t3[i]]."""
    lib = _lib.load()
    _need_cuda(f1.tensor, f2.tensor, f3.tensor, w)
    _bare(w, out, t1, t2, t3)
    P = t1.numel()
    assert t1.dtype == torch.int32 and t2.dtype == torch.int32 and t3.dtype == torch.int32
    n = P if rhs_only else P * P
    dev = f1.tensor.device
    if out is None:
        out = torch.empty((n,), dtype=torch.float64, device=dev)
        accumulate = False
    ks = lib.tn_generic_ksplit(rows, P, 1 if rhs_only else 0)
    work = torch.empty((ks * n,), dtype=torch.float64, device=dev) if (ks > 1 or accumulate) else None
    a, b, c = f1.c(), f2.c(), f3.c()
    rc = lib.tn_gram_generic(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.c_void_p(t1.data_ptr()),
                             ctypes.c_void_p(t2.data_ptr()), ctypes.c_void_p(t3.data_ptr()), P, _p(w), rows, _p(out),
                             1 if rhs_only else 0, _p(work), ks, 1 if accumulate else 0, _stream())
    _lib.check(rc, "tn_gram_generic")
    return out


def _int3(v):
    return (ctypes.c_int * 3)(*v)


@_op
def gram_sigma(M, m_pos, role_of_pos):
    lib = _lib.load()
    sigma = torch.empty((1,), dtype=torch.float64, device=M.device)
    _lib.check(lib.tn_gram_sigma(_p(M), _int3(m_pos), _int3(role_of_pos), _p(sigma), _stream()), "tn_gram_sigma")
    return sigma


@_op
def gram_expand(M, m_pos, role_of_pos, sigma, ridge, A=None):
    """Dense scaled system A = expand(M)/sigma + ridge*I, row stride padded to a multiple of 8."""
    lib = _lib.load()
    P = m_pos[0] * m_pos[1] * m_pos[2]
    lda = (P + 7) // 8 * 8
    if A is None:
        A = torch.empty((P, lda), dtype=torch.float64, device=M.device)
    elif A.dim() != 2 or A.shape[0] != P or A.stride(1) != 1 or A.stride(0) < P:
        raise _lib.TnError(f"gram_expand: A must be ({P}, >= {P}) with unit inner stride, got {tuple(A.shape)} / {A.stride()}")
    _need_cuda(M, sigma, A)
    _bare(M)
    _lib.check(lib.tn_gram_expand(_p(M), _int3(m_pos), _int3(role_of_pos), _p(sigma), float(ridge), _p(A), A.stride(0),
                                  _stream()), "tn_gram_expand")
    return A


@_op
def rhs_prepare(b, theta, sigma, ridge):
    lib = _lib.load()
    P = b.numel()
    out = torch.empty((P,), dtype=torch.float64, device=b.device)
    theta = theta.contiguous() if theta is not None else None
    _lib.check(lib.tn_rhs_prepare(_p(b), _p(theta), _p(sigma), float(ridge), _p(out), P, _stream()), "tn_rhs_prepare")
    return out


@_op
def cholesky_solve(A, rhs_vec):
    """In place: A (P x lda) <- its Cholesky factor (lower), rhs_vec <- solution.  Returns info tensor."""
    lib = _lib.load()
    P = A.shape[0]
    _need_cuda(A, rhs_vec)
    _system_layout(A, rhs_vec, P)
    work = torch.empty((lib.tn_cholesky_work_elems(P),), dtype=torch.float64, device=A.device)
    info = torch.zeros((1,), dtype=torch.int32, device=A.device)
    _lib.check(lib.tn_cholesky_solve(_p(A), A.stride(0), P, _p(rhs_vec), _p(work), ctypes.c_void_p(info.data_ptr()),
                                     _stream()), "tn_cholesky_solve")
    return info


@_op
def cholesky_solve_mixed(A, rhs_vec, rtol=1e-12, max_iter=12):
    """Tensor-core (3xTF32) factorisation + fp64 conjugate-gradient refinement.  A (P x lda, full symmetric) is overwritten
    below the diagonal, rhs_vec <- solution.  Returns (info, stats) device tensors; stats = [relative residual, iterations]."""
    lib = _lib.load()
    P = A.shape[0]
    _need_cuda(A, rhs_vec)
    _system_layout(A, rhs_vec, P)
    work = torch.empty((lib.tn_cholesky_mixed_work_elems(P),), dtype=torch.float64, device=A.device)
    info = torch.zeros((1,), dtype=torch.int32, device=A.device)
    stats = torch.zeros((2,), dtype=torch.float64, device=A.device)
    _lib.check(lib.tn_cholesky_solve_mixed(_p(A), A.stride(0), P, _p(rhs_vec), _p(work), ctypes.c_void_p(info.data_ptr()),
                                           float(rtol), int(max_iter), _p(stats), _stream()), "tn_cholesky_solve_mixed")
    return info, stats


@_op
def update_node(theta, step, lr=1.0, adaptive_step=False, max_norm=None):
    """theta (contiguous) <- theta + lr*step, in place."""
    lib = _lib.load()
    _need_cuda(theta, step)
    assert theta.is_contiguous() and step.is_contiguous()
    _lib.check(lib.tn_update_node(_p(theta), _p(step), theta.numel(), float(lr), 1 if adaptive_step else 0,
                                  float(max_norm) if max_norm is not None else -1.0, None, _stream()), "tn_update_node")
    return theta


@_op
def qr(a):
    """a (m x n contiguous, m >= n) <- Q in place; returns R (n x n)."""
    lib = _lib.load()
    _need_cuda(a)
    assert a.is_contiguous() and a.dim() == 2
    m, n = a.shape
    r = torch.empty((n, n), dtype=torch.float64, device=a.device)
    _lib.check(lib.tn_qr(_p(a), m, n, _p(r), _stream()), "tn_qr")
    return r


@_op
def matvec(fa: Factor, fb: Factor, fc: Factor, w, rows, v, out=None):
    """out = J^T diag(w) J v with J[row,(ia,ib,ic)] = fa fb fc."""
    lib = _lib.load()
    _need_cuda(fa.tensor, fb.tensor, fc.tensor, w, v)
    _bare(w, out)
    n = fa.m * fb.m * fc.m
    if out is None:
        out = torch.empty((n,), dtype=torch.float64, device=v.device)
    work = torch.empty((lib.tn_matvec_work_elems(rows, fa.m, fb.m, fc.m),), dtype=torch.float64, device=v.device)
    a, b, c = fa.c(), fb.c(), fc.c()
    v = v.contiguous()
    _lib.check(lib.tn_matvec_kr3(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), _p(w), rows, _p(v), _p(out), _p(work),
                                 _stream()), "tn_matvec_kr3")
    return out


@_op
def bmm(A, B, out=None, accumulate=False):
    """Per-sample matrix products out[s] (+)= A[s] @ B[s].  A: (S, I, K) or (I, K) shared; B: (S, K, J) or (K, J) shared;
    any strides (views are not copied).  Returns (S, I, J) contiguous."""
    lib = _lib.load()
    _need_cuda(A, B)
    S = A.shape[0] if A.dim() == 3 else B.shape[0]
    sA, (iA, kA) = (A.stride(0), A.stride()[1:]) if A.dim() == 3 else (0, A.stride())
    sB, (kB, jB) = (B.stride(0), B.stride()[1:]) if B.dim() == 3 else (0, B.stride())
    I, K = A.shape[-2:]
    K2, J = B.shape[-2:]
    assert K == K2, (A.shape, B.shape)
    if out is None:
        out = torch.empty((S, I, J), dtype=torch.float64, device=A.device)
        accumulate = False
    assert out.is_contiguous() and tuple(out.shape) == (S, I, J)
    _lib.check(lib.tn_bmm(_p(A), sA, iA, kA, _p(B), sB, kB, jB, _p(out), S, I, K, J, 1 if accumulate else 0, _stream()), "tn_bmm")
    return out


@_op
def outer_rows(G, W, w=None, gdiv=1, out=None, accumulate=False):
    """out (ra, m) (+)= sum_row w[row] * G[row // gdiv, :]^T W[row, :]; G (rows/gdiv, ra), W (rows, m), row strides free."""
    lib = _lib.load()
    _need_cuda(G, W, w)
    _bare(w)
    assert G.dim() == 2 and W.dim() == 2 and G.stride(1) == 1 and W.stride(1) == 1
    rows, m = W.shape
    ra = G.shape[1]
    if out is None:
        out = torch.empty((ra, m), dtype=torch.float64, device=W.device)
        accumulate = False
    assert out.is_contiguous()
    _lib.check(lib.tn_outer_rows(_p(G), G.stride(0), gdiv, ra, _p(W), W.stride(0), m, _p(w), rows, _p(out), 1 if accumulate else 0,
                                 _stream()), "tn_outer_rows")
    return out


@_op
def rows_dot(W, V, out=None):
    """z (rows, ra) = W (rows, m) @ V (ra, m)^T; row strides free, last dims contiguous.  ``out`` may be a 2-D view with its own row
    stride (a column block of a larger tensor)."""
    lib = _lib.load()
    _need_cuda(W, V)
    assert W.dim() == 2 and V.dim() == 2 and W.stride(1) == 1 and V.stride(1) == 1 and W.shape[1] == V.shape[1]
    rows, m = W.shape
    ra = V.shape[0]
    if out is None:
        out = torch.empty((rows, ra), dtype=torch.float64, device=W.device)
    assert out.dim() == 2 and tuple(out.shape) == (rows, ra) and (out.stride(1) == 1 or ra == 1)
    _lib.check(lib.tn_rows_dot(_p(W), W.stride(0), m, _p(V), V.stride(0), ra, rows, _p(out), out.stride(0) if rows > 1 else ra,
                               _stream()), "tn_rows_dot")
    return out


# ------------------------------------------------------------------------------------------ on-device Krylov drivers
class Operator:
    """The linear operator of a matrix-free local solve: ``Op v = A0 v / sigma + ridge v``.

    ``A0`` is ``J^T diag(w) J`` of this rank's rows given by three Kronecker ``factors`` (the built-in two-pass kernels), or an
    arbitrary ``matvec`` callable (conv-TT, cum-sum: their Jacobians are not Kronecker products); ``group`` sums ``A0 v`` over
    the ranks of a sample-sharded run.  Calling the object applies ``A0`` alone (plus the all-reduce) to a tensor -- the
    interface the SciPy bridge and tests use.
    """

    def __init__(self, P, factors=None, w=None, rows=0, matvec=None, group=None, sigma=None, ridge=0.0, device=None):
        if (factors is None) == (matvec is None):
            raise ValueError("Operator needs either three factors or a matvec callable")
        self.P = int(P)
        self.factors = factors
        self.w = w
        self.rows = int(rows)
        self.matvec = matvec
        self.group = group
        self.sigma = sigma
        self.ridge = float(ridge)
        self.device = device if device is not None else (factors[0].tensor.device if factors is not None else None)
        self.applies = 0            # operator applications served (bench bookkeeping)

    def with_shift(self, sigma, ridge):
        out = Operator(self.P, self.factors, self.w, self.rows, self.matvec, self.group, sigma, ridge, self.device)
        return out

    def __call__(self, v):
        self.applies += 1
        if self.matvec is not None:
            return self.matvec(v)
        fa, fb, fc = self.factors
        out = matvec(fa, fb, fc, self.w, self.rows, v.contiguous().view(-1))
        if self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(out, group=self.group)
        return out


class _OpBinding:
    """ctypes image of an Operator for the duration of one driver call: keeps the factor structs and callback thunks alive, maps
    the raw pointers the driver hands to the callbacks back to views of the tensors they point into, and carries a Python
    exception raised inside a callback across the C frame."""

    def __init__(self, op: Operator, buffers):
        self.op = op
        self.buffers = [t for t in buffers if t is not None]
        self.error = None
        self.keep = []
        c = _lib.tn_operator()
        if op.factors is not None:
            facs = [f.c() for f in op.factors]
            self.keep.append(facs)
            c.fa, c.fb, c.fc = (ctypes.pointer(f) for f in facs)
            c.w = 0 if op.w is None else op.w.data_ptr()
            c.rows = op.rows
            c.apply = 0
        else:
            cb = _lib.APPLY_FN(self._apply)
            self.keep.append(cb)
            c.apply = ctypes.cast(cb, ctypes.c_void_p).value
        c.apply_is_global = 1 if (op.group is not None and op.factors is None) else 0      # a python matvec does its own reduction
        if op.group is not None:
            cb = _lib.ALLREDUCE_FN(self._allreduce)
            self.keep.append(cb)
            c.allreduce = ctypes.cast(cb, ctypes.c_void_p).value
        else:
            c.allreduce = 0
        c.apply_ctx = c.allreduce_ctx = 0
        c.sigma = 0 if op.sigma is None else op.sigma.data_ptr()
        c.ridge = op.ridge
        c.P = op.P
        self.c = c

    def view(self, ptr, n):
        for t in self.buffers:
            base = t.data_ptr()
            if base <= ptr and ptr + 8 * n <= base + 8 * t.numel():
                off = (ptr - base) // 8
                return t.view(-1)[off:off + n]
        raise _lib.TnError("krylov callback: pointer outside the buffers of the call")

    def _apply(self, ctx, v_ptr, out_ptr, stream):
        try:
            v = self.view(v_ptr, self.op.P)
            out = self.view(out_ptr, self.op.P)
            self.op.applies += 1
            out.copy_(self.op.matvec(v).reshape(-1))
            return 0
        except BaseException as e:       # must not propagate through the C frame
            self.error = e
            return -1

    def _allreduce(self, ctx, buf_ptr, n, stream):
        try:
            import torch.distributed as dist
            dist.all_reduce(self.view(buf_ptr, n), group=self.op.group)
            return 0
        except BaseException as e:
            self.error = e
            return -1

    def finish(self, rc, what):
        if self.error is not None:
            raise self.error
        _lib.check(rc, what)


def _krylov_common(op: Operator, b, x0):
    _need_cuda(b, x0, op.sigma, op.w)
    _bare(b, x0, op.w)
    if b.numel() != op.P or (x0 is not None and x0.numel() != op.P):
        raise _lib.TnError(f"krylov: vectors of {op.P} expected")
    return b.device


def _poll(op, poll_every):
    if poll_every is not None:
        return int(poll_every)
    # every iteration: the operator's kernels do not look at the convergence flag, so an unpolled iteration after convergence is a
    # wasted pass over the rows (measured on config 4a: 3x the time of the whole site update), a poll is one stream sync
    return 1


@_op
def cholesky_factor(A, tensor_core=False):
    """In place: lower triangle of A (P x lda) <- its Cholesky factor.  Returns (work, info) for cholesky_apply / cg.
    tensor_core: False / 0 = fp64, True / 1 = 3xTF32 trailing updates, 2 = one TF32 pass (a factor that only preconditions)."""
    lib = _lib.load()
    P = A.shape[0]
    _need_cuda(A)
    if A.dim() != 2 or A.stride(1) != 1 or A.stride(0) < P:
        raise _lib.TnError(f"system matrix must be ({P}, >= {P}) with unit inner stride, got {tuple(A.shape)} / {A.stride()}")
    work = torch.empty((lib.tn_cholesky_work_elems(P),), dtype=torch.float64, device=A.device)
    info = torch.zeros((1,), dtype=torch.int32, device=A.device)
    _lib.check(lib.tn_cholesky_factor(_p(A), A.stride(0), P, int(tensor_core), _p(work), ctypes.c_void_p(info.data_ptr()),
                                      _stream()), "tn_cholesky_factor")
    return work, info


@_op
def cholesky_apply(L, work, info, x):
    """x <- L^-T L^-1 x in place."""
    lib = _lib.load()
    P = L.shape[0]
    _need_cuda(L, work, x)
    _system_layout(L, x, P)
    _lib.check(lib.tn_cholesky_apply(_p(L), L.stride(0), P, _p(x), _p(work), ctypes.c_void_p(info.data_ptr()), _stream()),
               "tn_cholesky_apply")
    return x


@_op
def gram_trace(fa: Factor, fb: Factor, fc: Factor, w, rows, out=None):
    """out[0] = sum_rows w |fa|^2 |fb|^2 |fc|^2 (the trace of the local Gram, fp64), out[1] the same with |w|."""
    lib = _lib.load()
    _need_cuda(fa.tensor, fb.tensor, fc.tensor, w, out)
    _bare(w, out)
    if out is None:
        out = torch.empty((2,), dtype=torch.float64, device=fa.tensor.device)
    a, b, c = fa.c(), fb.c(), fc.c()
    _lib.check(lib.tn_gram_trace(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), _p(w), rows, _p(out), 0, _stream()), "tn_gram_trace")
    return out


def cg(op: Operator, b, x0=None, precond=None, max_iter=50, rtol=1e-6, poll_every=None):
    """Conjugate gradients on Op x = b on the device.  ``precond = (L, work, info)`` from cholesky_factor.  Stops at
    |r| <= rtol |b|, or -- preconditioned -- at |L^-T L^-1 r| <= rtol |x| (an estimate of the relative forward error).
    Returns (x, stats) with stats = [relative residual, iterations, stopped-by-tolerance, operator applications, value of the
    stopping criterion] (device)."""
    lib = _lib.load()
    dev = _krylov_common(op, b, x0)
    with _on(dev):
        x = torch.empty_like(b) if x0 is None else x0.clone().reshape(-1)
        stats = torch.zeros((5,), dtype=torch.float64, device=dev)
        bind = _OpBinding(op, [b, x])
        work = torch.empty((lib.tn_cg_work_elems(ctypes.byref(bind.c)),), dtype=torch.float64, device=dev)
        bind.buffers.append(work)
        L, lwork, linfo = precond if precond is not None else (None, None, None)
        if L is not None:
            _need_cuda(L, lwork)
            _system_layout(L, b, op.P)
        rc = lib.tn_cg(ctypes.byref(bind.c), _p(L), 0 if L is None else L.stride(0), _p(lwork),
                       ctypes.c_void_p(0 if linfo is None else linfo.data_ptr()), _p(b), _p(x), 0 if x0 is None else 1, int(max_iter),
                       float(rtol), _poll(op, poll_every), _p(work), _p(stats), _stream())
        bind.finish(rc, "tn_cg")
    return x, stats


def minres(op: Operator, b, x0=None, max_iter=50, rtol=1e-6, poll_every=None):
    """MINRES on Op x = b on the device; returns (x, stats) as cg."""
    lib = _lib.load()
    dev = _krylov_common(op, b, x0)
    with _on(dev):
        x = torch.empty_like(b) if x0 is None else x0.clone().reshape(-1)
        stats = torch.zeros((4,), dtype=torch.float64, device=dev)
        bind = _OpBinding(op, [b, x])
        work = torch.empty((lib.tn_minres_work_elems(ctypes.byref(bind.c)),), dtype=torch.float64, device=dev)
        bind.buffers.append(work)
        rc = lib.tn_minres(ctypes.byref(bind.c), _p(b), _p(x), 0 if x0 is None else 1, int(max_iter), float(rtol), _poll(op, poll_every),
                           _p(work), _p(stats), _stream())
        bind.finish(rc, "tn_minres")
    return x, stats


def lanczos(op: Operator, b, x0=None, max_iter=50, tol=1e-6, poll_every=None):
    """Lanczos-Galerkin solve of Op x = b started at x0 (reference network.py:793-824); returns (x, stats)."""
    lib = _lib.load()
    dev = _krylov_common(op, b, x0)
    with _on(dev):
        x = torch.empty_like(b)
        x0c = None if x0 is None else x0.contiguous().reshape(-1)
        stats = torch.zeros((4,), dtype=torch.float64, device=dev)
        bind = _OpBinding(op, [b, x, x0c])
        work = torch.empty((lib.tn_lanczos_work_elems(ctypes.byref(bind.c), int(max_iter)),), dtype=torch.float64, device=dev)
        bind.buffers.append(work)
        rc = lib.tn_lanczos(ctypes.byref(bind.c), _p(b), _p(x0c), _p(x), int(max_iter), float(tol), _poll(op, poll_every), _p(work),
                            _p(stats), _stream())
        bind.finish(rc, "tn_lanczos")
    return x, stats
