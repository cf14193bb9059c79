"""Offline (CPU) experiment: how good a preconditioner is a Gram whose operand tiles are rounded to TF32 / FP16 / BF16?

Builds a config-5a-like middle site at reduced rank (P = r*29*r), the exact fp64 system (A/sigma + ridge I), and emulated tensor-core
Grams: operand entries (w*pair(fa)*pair(fb) and pair(fc)) rounded to the format, exact products, fp32 accumulation.  Reports the
spectrum of the preconditioned operator and the PCG iterations to 1e-11.       python tools/precond_experiment.py [r] [rows]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_default_dtype(torch.float64)
r = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
F = 28
rng = np.random.default_rng(0)
X = np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1)
w1 = rng.normal(size=(F + 1, 1)) / F ** 0.5; w2 = rng.normal(size=(F + 1, 1)) / F ** 0.5
y = np.tanh(X @ w1) + 0.5 * (X @ w2) ** 2 + 0.1 * rng.normal(size=(N, 1))
import tensornetworksfork_b200 as tnb
from oracle import tn_oracle as orc
layer = tnb.TensorTrainLayer(5, r, F + 1, output_shape=1, constrict_bond=False, perturb=False, seed=42)
cores = [n.tensor.numpy().copy() for n in layer.tensor_network.train_nodes]
t0 = time.time()
sweeps = int(os.environ.get("SWEEPS", "1"))
if sweeps:
    orc.accumulating_swipe(cores, X, y, loss="square", batch_size=-1, num_swipes=sweeps, method="ridge_cholesky", eps=1.0)
print("sweeps done", time.time() - t0, flush=True)
phis = orc.site_inputs(X, 5)
Ls, Rs = orc.left_envs(cores, phis), orc.right_envs(cores, phis)
k = 2
fa, fb, fc = Ls[k - 1][:, 0, :], X, Rs[k + 1][:, 0, :]      # (N, r), (N, 29), (N, r)
print("factor magnitudes", np.abs(fa).max(), np.abs(fb).max(), np.abs(fc).max(), "min|fa|", np.abs(fa).min(), np.abs(fc).min())
wts = np.full(N, 2.0)
J = np.einsum("sa,sp,sb->sapb", fa, fb, fc).reshape(N, -1)
P = J.shape[1]
A = (J * wts[:, None]).T @ J
sigma = np.abs(np.diag(A)).mean()

def rnd(x, fmt):
    x = np.asarray(x, dtype=np.float32)
    if fmt == "tf32":
        u = x.view(np.uint32); u = (u + 0x1000) & 0xffffe000; return u.view(np.float32)
    if fmt == "fp16":
        return x.astype(np.float16).astype(np.float32)
    if fmt == "bf16":
        u = x.view(np.uint32); u = (u + 0x8000) & 0xffff0000; return u.view(np.float32)
    return x

def pow2scale(v):
    e = np.ceil(np.log2(np.abs(v).max())); return v / 2.0 ** e, 2.0 ** e

def approx_gram(fmt, half_math):
    a, sa = pow2scale(fa); b, sb = pow2scale(fb); c, sc = pow2scale(fc); ww, sw = pow2scale(wts)
    ia, ja = np.triu_indices(r); ib, jb = np.triu_indices(F + 1)
    if half_math:    # factors stored in the format, every product rounded to the format (packed half math in the producers)
        a32, b32, c32 = rnd(a, fmt), rnd(b, fmt), rnd(c, fmt)
        wa = rnd(rnd(ww[:, None].astype(np.float32), fmt) * a32, fmt)
        PA = rnd(wa[:, ia] * a32[:, ja], fmt); PB = rnd(b32[:, ib] * b32[:, jb], fmt)
        V = rnd(c32[:, ia] * c32[:, ja], fmt)
        U = rnd(PA[:, :, None] * PB[:, None, :], fmt).reshape(N, -1)
    else:            # fp32 products, one rounding of the tile entry
        a32, b32, c32 = a.astype(np.float32), b.astype(np.float32), c.astype(np.float32)
        PA = (ww[:, None].astype(np.float32) * a32)[:, ia] * a32[:, ja]; PB = b32[:, ib] * b32[:, jb]
        V = rnd(c32[:, ia] * c32[:, ja], fmt)
        U = rnd((PA[:, :, None] * PB[:, None, :]).reshape(N, -1), fmt)
    M = (U.T @ V).astype(np.float64) * (sa * sa * sb * sb * sc * sc * sw)          # fp32 accumulation (BLAS), unscale
    # expand unique entries to dense A: index maps
    na, nb = len(ia), len(ib)
    qa = np.zeros((r, r), int); qa[ia, ja] = np.arange(na); qa[ja, ia] = np.arange(na)
    qb = np.zeros((F + 1, F + 1), int); qb[ib, jb] = np.arange(nb); qb[jb, ib] = np.arange(nb)
    M3 = M.reshape(na, nb, na)
    idx = np.arange(P); i_a = idx // ((F + 1) * r); i_p = (idx // r) % (F + 1); i_b = idx % r
    return M3[qa[i_a[:, None], i_a[None, :]], qb[i_p[:, None], i_p[None, :]], qa[i_b[:, None], i_b[None, :]]]

def pcg_iters(Aex, Aap, ridge, tol=1e-11, maxit=200):
    Pn = Aex.shape[0]
    Op = Aex / sigma + ridge * np.eye(Pn)
    Mp = Aap / sigma + ridge * np.eye(Pn)
    try:
        L = np.linalg.cholesky(Mp)
    except np.linalg.LinAlgError:
        return None, None
    import scipy.linalg as sl
    b = np.random.default_rng(1).normal(size=Pn)
    prec = lambda v: sl.cho_solve((L, True), v)
    x = prec(b); rres = b - Op @ x; z = prec(rres); p = z.copy(); rz = rres @ z
    for it in range(maxit):
        if np.linalg.norm(z) <= tol * np.linalg.norm(x):
            return it, np.linalg.norm(np.linalg.solve(Op, b) - x) / np.linalg.norm(x)
        q = Op @ p; al = rz / (p @ q); x += al * p; rres -= al * q; z = prec(rres); rz2 = rres @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit, None

ev = np.linalg.eigvalsh(A / sigma)
print(f"P={P} sigma={sigma:.3e} lambda_max/sigma={ev[-1]:.3e} lambda_min/sigma={ev[0]:.3e}", flush=True)
for fmt, hm in (("tf32", False), ("fp16", False), ("fp16", True), ("bf16", False), ("bf16", True)):
    Aap = approx_gram(fmt, hm)
    err = np.linalg.norm(Aap - A) / np.linalg.norm(A)
    for ridge in (2.0, 0.25, 2e-2, 2e-3):
        it, fe = pcg_iters(A, Aap, ridge)
        print(f"{fmt:5s} half_math={hm!s:5s} rel_fro_err={err:.2e} ridge={ridge:g}: pcg iterations={it} fwd_err={fe}", flush=True)
