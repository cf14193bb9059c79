"""Kernel-level parity through the C ABI against the numpy oracle (needs a B200)."""
import math

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import tn_oracle as orc

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)

from tensornetworksfork_b200 import ops  # noqa: E402
from tensornetworksfork_b200.ops import Factor  # noqa: E402

DEV = "cuda"


def T(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


def pairs(F):
    m = F.shape[1]
    iu = [(i, j) for i in range(m) for j in range(i, m)]
    return np.stack([F[:, i] * F[:, j] for i, j in iu], axis=1)


@pytest.mark.parametrize("S,rin,f,rout", [(1, 1, 3, 2), (200, 4, 5, 6), (1000, 24, 2, 24), (333, 38, 29, 38), (130, 3, 7, 100),
                                          (64, 100, 9, 70), (4177, 6, 9, 6)])
def test_env_update_identity(S, rin, f, rout):
    rng = np.random.default_rng(S + rin)
    env = rng.normal(size=(S, rin))
    x = rng.uniform(-1, 1, size=(S, f))
    core = rng.normal(size=(rin, f, rout))
    want = np.einsum("sa,sp,apb->sb", env, x, core)
    got = ops.env_update(T(env), Factor(T(x), m=f), T(core), S).cpu().numpy()
    assert gu.relerr(got, want) < 1e-13
    dot = rng.normal(size=(S, rout))
    yh = ops.predict(T(env), Factor(T(x), m=f), T(core), T(dot), S).cpu().numpy()
    assert gu.relerr(yh, (want * dot).sum(1)) < 1e-13


@pytest.mark.parametrize("S,C,rin,f,rout", [(120000, 1, 24, 2, 24), (90001, 3, 6, 3, 5), (100000, 1, 1, 4, 7)])
def test_env_update_persistent_pipelined_path(S, C, rin, f, rout):
    """Enough row tiles to take the persistent cp.async-pipelined DMMA kernel (small contraction, core resident)."""
    rng = np.random.default_rng(S)
    env = rng.normal(size=(S * C, rin))
    x = rng.uniform(-1, 1, size=(S, f))
    core = rng.normal(size=(rin, f, rout))
    dot = rng.normal(size=(S, rout))
    xr = np.repeat(x, C, axis=0)
    want = np.einsum("sa,sp,apb->sb", env, xr, core)
    e = None if rin == 1 else T(env)
    if rin == 1:
        want = np.einsum("sp,pb->sb", xr, core[0])
    got = ops.env_update(e, Factor(T(x), m=f), T(core), S * C, cdiv=C).cpu().numpy()
    assert gu.relerr(got, want) < 1e-13
    yh = ops.predict(e, Factor(T(x), m=f), T(core), T(dot), S * C, cdiv=C, dot_div=C).cpu().numpy()
    assert gu.relerr(yh, (want * np.repeat(dot, C, axis=0)).sum(1)) < 1e-13


def test_env_update_chain_end_and_class_rows():
    rng = np.random.default_rng(1)
    S, C, f, r = 300, 3, 4, 5
    x = rng.uniform(-1, 1, size=(S, f))
    core = rng.normal(size=(1, f, r))
    got = ops.env_update(None, Factor(T(x), m=f), T(core), S).cpu().numpy()
    assert gu.relerr(got, x @ core[0]) < 1e-14
    env = rng.normal(size=(S, C, r))
    core2 = rng.normal(size=(r, f, 6))
    got = ops.env_update(T(env).reshape(S * C, r), Factor(T(x), m=f), T(core2), S * C, cdiv=C).cpu().numpy()
    want = np.einsum("sca,sp,apb->scb", env, x, core2).reshape(S * C, 6)
    assert gu.relerr(got, want) < 1e-13
    U = rng.normal(size=(S, 4, C))
    g = rng.normal(size=(S, C))
    F, G = ops.class_rows(T(env), T(U), T(g))
    assert gu.relerr(F.cpu().numpy(), np.einsum("stc,sca->sta", U, env).reshape(S * 4, r)) < 1e-13
    assert gu.relerr(G.cpu().numpy(), np.einsum("sc,sca->sa", g, env)) < 1e-13


def test_env_update_feature_maps():
    rng = np.random.default_rng(2)
    S, F_, r = 500, 6, 7
    X = rng.uniform(-1, 1, size=(S, F_))
    env = rng.normal(size=(S, r))
    for kind, f, mk, phi in (("sincos", 2, ops.MAP_SINCOS, orc.fbasis(X)), ("poly", 4, ops.MAP_POLY, orc.polynomial_basis(X, 3))):
        core = rng.normal(size=(r, f, 5))
        for col in (0, 3, F_ - 1):
            want = np.einsum("sa,sp,apb->sb", env, phi[col], core)
            got = ops.env_update(T(env), Factor(T(X), m=f, map_kind=mk, col=col), T(core), S).cpu().numpy()
            assert gu.relerr(got, want) < 1e-13, (kind, col)


@pytest.mark.parametrize("S,ma,mb,mc,V", [(50, 1, 3, 2, 1), (777, 4, 5, 4, 1), (1500, 6, 9, 6, 1), (400, 3, 4, 1, 1), (300, 5, 3, 4, 3),
                                          (5000, 24, 2, 24, 1), (260, 12, 7, 12, 1)])
def test_gram_and_rhs_fp64(S, ma, mb, mc, V):
    rng = np.random.default_rng(S)
    rows = S * V
    Fa = rng.normal(size=(rows, ma))
    Fb = rng.uniform(-1, 1, size=(S, mb))
    Fc = rng.normal(size=(S, mc))
    w = rng.normal(size=rows)
    Fb_r, Fc_r = np.repeat(Fb, V, axis=0), np.repeat(Fc, V, axis=0)
    J = np.einsum("ra,rp,rb->rapb", Fa, Fb_r, Fc_r).reshape(rows, -1)
    A_want = (J * w[:, None]).T @ J
    b_want = J.T @ w
    fa, fb, fc = Factor(T(Fa), m=ma), Factor(T(Fb), m=mb, div=V), Factor(T(Fc), m=mc, div=V)
    for role_of_pos, order in (([0, 1, 2], (0, 1, 2)), ([2, 0, 1], (1, 2, 0)), ([0, 2, 1], (0, 2, 1))):
        facs = (fa, fb, fc)
        M = ops.gram(ops.GRAM_FP64, facs[order[0]], facs[order[1]], facs[order[2]], T(w), rows)
        one = torch.ones(1, device=DEV)
        A = ops.gram_expand(M, (ma, mb, mc), role_of_pos, one, 0.0)
        P = ma * mb * mc
        assert gu.relerr(A[:, :P].cpu().numpy(), A_want) < 1e-12, role_of_pos
        sig = ops.gram_sigma(M, (ma, mb, mc), role_of_pos)
        assert abs(float(sig) - np.abs(np.diag(A_want)).mean()) < 1e-12 * max(1.0, np.abs(np.diag(A_want)).mean())
    b = ops.rhs(fa, fb, fc, T(w), rows).cpu().numpy()
    assert gu.relerr(b, b_want) < 1e-12
    M2 = ops.gram(ops.GRAM_FP64, fa, fb, fc, T(w), rows)
    M3 = ops.gram(ops.GRAM_FP64, fa, fb, fc, T(w), rows, M=M2.clone(), accumulate=True)
    assert gu.relerr(M3.cpu().numpy(), 2 * M2.cpu().numpy()) < 1e-14
    v = rng.normal(size=ma * mb * mc)
    mv = ops.matvec(fa, fb, fc, T(w), rows, T(v)).cpu().numpy()
    assert gu.relerr(mv, A_want @ v) < 1e-11


@pytest.mark.parametrize("P", [1, 5, 64, 65, 100, 324, 900, 1200, 2888, 4500, 6001])
def test_cholesky_solve(P):
    rng = np.random.default_rng(P)
    B = rng.normal(size=(P, P + 10))
    A = B @ B.T / P + 0.5 * np.eye(P)
    rhs = rng.normal(size=P)
    lda = (P + 7) // 8 * 8
    Ap = torch.zeros((P, lda), device=DEV)
    Ap[:, :P] = T(A)
    r = T(rhs)
    info = ops.cholesky_solve(Ap, r)
    assert int(info.item()) == 0
    x = r.cpu().numpy()
    res = np.linalg.norm(A @ x - rhs) / np.linalg.norm(rhs)
    assert res < 1e-11, res
    Lref = np.linalg.cholesky(A)
    Lgot = np.tril(Ap[:, :P].cpu().numpy())
    assert gu.relerr(Lgot, Lref) < 1e-11


def test_cholesky_reports_non_spd():
    P = 150
    rng = np.random.default_rng(0)
    B = rng.normal(size=(P, P))
    A = B @ B.T
    A[100, 100] = -1.0
    Ap = torch.zeros((P, 152), device=DEV)
    Ap[:, :P] = T(A)
    r = T(rng.normal(size=P))
    info = ops.cholesky_solve(Ap, r)
    assert int(info.item()) == 101


def test_solve_pipeline_matches_oracle():
    rng = np.random.default_rng(3)
    S, ma, mb, mc = 900, 4, 5, 4
    Fa, Fb, Fc = rng.normal(size=(S, ma)), rng.uniform(-1, 1, size=(S, mb)), rng.normal(size=(S, mc))
    w = np.full(S, 2.0)
    g = rng.normal(size=S)
    theta = rng.normal(size=ma * mb * mc)
    J = np.einsum("ra,rp,rb->rapb", Fa, Fb, Fc).reshape(S, 1, -1)
    A, b = orc.gram(J, g[:, None], w[:, None, None])
    for eps in (1.0, 1e-3):
        want = orc.solve_system(A, b, theta, "ridge_cholesky", eps)
        fa, fb, fc = Factor(T(Fa), m=ma), Factor(T(Fb), m=mb), Factor(T(Fc), m=mc)
        M = ops.gram(ops.GRAM_FP64, fa, fb, fc, T(w), S)
        bb = ops.rhs(fa, fb, fc, T(g), S)
        pos = (ma, mb, mc)
        sig = ops.gram_sigma(M, pos, [0, 1, 2])
        Ad = ops.gram_expand(M, pos, [0, 1, 2], sig, 2 * eps)
        rv = ops.rhs_prepare(bb, T(theta), sig, 2 * eps)
        info = ops.cholesky_solve(Ad, rv)
        assert int(info.item()) == 0
        assert gu.relerr(rv.cpu().numpy(), want) < 1e-10


def test_update_node_variants():
    rng = np.random.default_rng(4)
    th, st = rng.normal(size=1000), 5 * rng.normal(size=1000)
    for kw in (dict(lr=1.0), dict(lr=0.3), dict(lr=1.0, adaptive_step=True), dict(lr=1.0, max_norm=2.0),
               dict(lr=0.5, adaptive_step=True, max_norm=1.0)):
        want = orc.update_node(th, st, **kw)
        t = T(th)
        ops.update_node(t, T(st), **kw)
        assert gu.relerr(t.cpu().numpy(), want) < 1e-14, kw


@pytest.mark.parametrize("m,n", [(4, 2), (8, 4), (48, 24), (76, 38), (228, 38), (100, 100), (1102, 38), (9, 1)])
def test_qr_matches_lapack_convention(m, n):
    rng = np.random.default_rng(m * n)
    A = rng.normal(size=(m, n))
    Q, R = np.linalg.qr(A, mode="reduced")
    a = T(A)
    r = ops.qr(a)
    assert gu.relerr(r.cpu().numpy(), R) < 1e-12
    assert gu.relerr(a.cpu().numpy(), Q) < 1e-12


def test_env_update_full_size_linearity():
    """Config-3 size (N=515k, r=24, f=2): linearity in the environment and agreement of a strided
    sample of rows with the oracle -- properties that do not need the oracle at full size."""
    S, r, f = 515345, 24, 2
    g = torch.Generator(device=DEV).manual_seed(0)
    e1 = torch.randn((S, r), device=DEV, generator=g)
    e2 = torch.randn((S, r), device=DEV, generator=g)
    X = torch.rand((S, 90), device=DEV, generator=g) * 2 - 1
    core = torch.randn((r, f, r), device=DEV, generator=g)
    fx = Factor(X, m=2, map_kind=ops.MAP_SINCOS, col=17)
    o1 = ops.env_update(e1, fx, core, S)
    o2 = ops.env_update(e2, fx, core, S)
    o12 = ops.env_update(e1 + 2 * e2, fx, core, S)
    assert float((o12 - (o1 + 2 * o2)).abs().max()) < 1e-12 * float(o12.abs().max())
    idx = torch.arange(0, S, 5003, device=DEV)
    phi = orc.fbasis(X[idx].cpu().numpy())[17]
    want = np.einsum("sa,sp,apb->sb", e1[idx].cpu().numpy(), phi, core.cpu().numpy())
    assert gu.relerr(o1[idx].cpu().numpy(), want) < 1e-13


@pytest.mark.parametrize("S,C,rin,f,rout,fmap", [(200003, 1, 24, 2, 24, "sincos"), (160000, 1, 38, 2, 38, "sincos"), (40000, 5, 16, 2, 12, "sincos"),
                                                 (170000, 1, 1, 4, 20, None), (180000, 1, 12, 3, 10, None), (150016, 1, 8, 4, 6, "poly")])
def test_env_update_warp_per_tile_bulk_copy_path(S, C, rin, f, rout, fmap):
    """Rows enough for the warp-per-tile kernel (env.cu::env_warp_kernel: every warp owns 16-row tiles, environment rows fetched
    and results stored with bulk copies, no block barrier in the loop): dense environments, feature maps, a ragged last tile, the
    prediction epilogue, class rows sharing a site input (cdiv), a strided output; against numpy on a sample of rows."""
    rng = np.random.default_rng(S % 1000 + rin)
    rows = S * C
    env = rng.normal(size=(rows, rin))
    X = rng.uniform(-1, 1, size=(S, 6))
    core = rng.normal(size=(rin, f, rout))
    if fmap is None:
        fx = Factor(T(X[:, :f].copy()), m=f)
        phi = X[:, :f]
    else:
        mk = ops.MAP_SINCOS if fmap == "sincos" else ops.MAP_POLY
        fx = Factor(T(X), m=f, map_kind=mk, col=2)
        phi = np.stack([np.cos(0.5 * np.pi * X[:, 2]), np.sin(0.5 * np.pi * X[:, 2])], 1) if fmap == "sincos" else np.stack([X[:, 2] ** d for d in range(f)], 1)
    e = None if rin == 1 else T(env)
    got = ops.env_update(e, fx, T(core), rows, cdiv=C).cpu().numpy()
    idx = np.unique(np.concatenate([np.arange(0, rows, 997), np.arange(rows - 40, rows), np.arange(0, 40)]))
    ein = np.ones((len(idx), 1)) if rin == 1 else env[idx]
    want = np.einsum("sa,sp,apb->sb", ein, phi[idx // C], core)
    assert gu.relerr(got[idx], want) < 1e-13
    dot = rng.normal(size=(S, rout))
    yh = ops.predict(e, fx, T(core), T(dot), rows, cdiv=C, dot_div=C).cpu().numpy()
    assert gu.relerr(yh[idx], (want * dot[idx // C]).sum(1)) < 1e-13
    wide = torch.zeros((rows, rout + 3), dtype=torch.float64, device="cuda")      # a column block of a wider tensor: row stride != r_out
    ops.env_update(e, fx, T(core), rows, cdiv=C, out=wide[:, 1:1 + rout])
    assert gu.relerr(wide[:, 1:1 + rout].cpu().numpy()[idx], want) < 1e-13 and float(wide[:, 0].abs().max()) == 0.0 and float(wide[:, 1 + rout:].abs().max()) == 0.0
    if e is not None:       # environment rows as a column block of a wider tensor (row stride != r_in, still 16-byte aligned rows)
        big = torch.zeros((rows, rin + 4), dtype=torch.float64, device="cuda")
        big[:, 2:2 + rin] = e
        got2 = ops.env_update(big[:, 2:2 + rin], fx, T(core), rows, cdiv=C)
        assert torch.equal(got2, torch.as_tensor(got, device="cuda"))


@pytest.mark.parametrize("S,ma,mb,mc,V,fmap", [(20011, 38, 29, 38, 1, None), (9000, 38, 6, 38, 1, "poly"), (4100, 24, 12, 17, 1, None),
                                              (3000, 40, 3, 40, 1, None), (2500, 9, 60, 8, 1, None), (1300, 38, 4, 38, 3, "poly"), (37, 33, 5, 33, 1, None)])
def test_rhs_large_core_register_blocked(S, ma, mb, mc, V, fmap, monkeypatch):
    """Right-hand side / J^T u for cores with P > 4096 and outer factors <= 40 wide (gram.cu::rhs_big_kernel: a warp owns one
    middle index and its whole output block in registers, the middle factor enters as a per-row scalar): against numpy and
    against the GEMM-shaped kernel it replaces; ragged row counts, a mapped middle factor, class rows sharing a site input,
    accumulation into an existing vector, fewer rows than one split."""
    rng = np.random.default_rng(S + ma)
    rows = S * V
    Fa = rng.normal(size=(rows, ma))
    X = rng.uniform(-1, 1, size=(S, max(mb, 4)))
    Fc = rng.normal(size=(S, mc))
    w = rng.normal(size=rows)
    if fmap is None:
        fb = Factor(T(X[:, :mb].copy()), m=mb, div=V)
        phi = X[:, :mb]
    elif fmap == "poly":
        fb = Factor(T(X), m=mb, map_kind=ops.MAP_POLY, col=1, div=V)
        phi = np.stack([X[:, 1] ** d for d in range(mb)], 1)
    else:
        fb = Factor(T(X), m=mb, map_kind=ops.MAP_SINCOS, col=2, div=V)
        phi = np.stack([np.cos(0.5 * np.pi * X[:, 2]), np.sin(0.5 * np.pi * X[:, 2])], 1)
    fa, fc = Factor(T(Fa), m=ma), Factor(T(Fc), m=mc, div=V)
    want = np.einsum("r,ra,rp,rb->apb", w, Fa, np.repeat(phi, V, axis=0), np.repeat(Fc, V, axis=0)).reshape(-1)
    got = ops.rhs(fa, fb, fc, T(w), rows)
    assert gu.relerr(got.cpu().numpy(), want) < 1e-12
    twice = ops.rhs(fa, fb, fc, T(w), rows, b=got.clone(), accumulate=True)
    assert gu.relerr(twice.cpu().numpy(), 2 * want) < 1e-12
    monkeypatch.setenv("TN_RHS_NO_BIG", "1")
    old = ops.rhs(fa, fb, fc, T(w), rows)
    assert gu.relerr(old.cpu().numpy(), want) < 1e-12


@pytest.mark.parametrize("S,ma,mb,mc,opts", [
    (37, 3, 2, 4, {}),                                  # fewer rows than two chunks
    (1000, 5, 12, 7, {}),                               # 16 x 8 box tiling, ragged boxes on both sides
    (333, 6, 3, 6, {"no_w": True}),                     # consecutive-index tiling, unit weights
    (2051, 10, 2, 9, {"map_b": "sincos"}),              # mapped middle factor expanded by the two mapper warps
    (4100, 4, 11, 5, {"map_b": "poly"}),                # mapped factor on the box tiling
    (640, 5, 3, 6, {"div_a": 4}),                       # rows of the first factor shared by 4 samples
    (900, 17, 1, 13, {}),                               # one pair in the middle: the unfactored kernel keeps the site
    (3000, 2, 24, 24, {"map_a": "sincos"}),             # the engine's role order of config 3: first two factors exchanged
    (1200, 6, 38, 38, {}),                              # ... of config 5b
    (800, 11, 12, 3, {}),                               # exchanged factors on the box tiling
    (70000, 7, 2, 7, {"accumulate": True}),             # several row splits reduced into an existing M
    (300, 40, 12, 38, {}),                              # wide factors (two cp.async rounds per row)
])
def test_gram_fp64_factored_operand_kernel(S, ma, mb, mc, opts, monkeypatch):
    """fp64 Gram with the factored left operand (gram.cu::gram_f64_fact_kernel: the tile holds [w pair(fa) | pair(fb) | pair(fc)] and
    an A-fragment element is one product made in registers; cp.async ring of raw factors; one barrier per 32-row chunk) on every
    tiling the launcher can choose, against numpy and against the unfactored kernel (TN_GRAM_F64_UNFACTORED=1)."""
    rng = np.random.default_rng(S + 7 * ma + mb)
    da = opts.get("div_a", 1)

    def feat(raw, kind, m):
        if kind == "sincos":
            return np.stack([np.cos(0.5 * np.pi * raw), np.sin(0.5 * np.pi * raw)], 1)
        return np.stack([raw ** d for d in range(m)], 1)

    kinds = {"sincos": ops.MAP_SINCOS, "poly": ops.MAP_POLY}
    if "map_a" in opts:
        Xa = rng.uniform(-1, 1, size=(S, 3))
        fa, Fa = Factor(T(Xa), m=ma, map_kind=kinds[opts["map_a"]], col=2), feat(Xa[:, 2], opts["map_a"], ma)
    else:
        Ta = rng.normal(size=(S // da, ma))
        fa, Fa = Factor(T(Ta), m=ma, div=da), np.repeat(Ta, da, axis=0)
    if "map_b" in opts:
        Xb = rng.uniform(-1, 1, size=(S, 4))
        fb, Fb = Factor(T(Xb), m=mb, map_kind=kinds[opts["map_b"]], col=1), feat(Xb[:, 1], opts["map_b"], mb)
    else:
        Fb = rng.uniform(-1, 1, size=(S, mb))
        fb = Factor(T(Fb), m=mb)
    Fc_wide = rng.normal(size=(S, mc + 3))
    Fc = Fc_wide[:, :mc]
    fc = Factor(T(Fc_wide)[:, :mc], m=mc)              # row stride != m
    w = None if opts.get("no_w") else rng.uniform(0.5, 1.5, size=S)
    U = (pairs(Fa)[:, :, None] * pairs(Fb)[:, None, :]).reshape(S, -1) * (np.ones(S) if w is None else w)[:, None]
    want = (U.T @ pairs(Fc)).reshape(-1)                 # M = U^T V as one BLAS product (an einsum over four operands takes minutes here)
    wt = None if w is None else T(w)

    def run():
        if opts.get("accumulate"):
            M0 = torch.full((want.size,), 0.25, device=DEV)
            return (ops.gram(ops.GRAM_FP64, fa, fb, fc, wt, S, M=M0, accumulate=True) - 0.25).cpu().numpy()
        return ops.gram(ops.GRAM_FP64, fa, fb, fc, wt, S).cpu().numpy()

    got = run()
    assert np.isfinite(got).all()
    assert gu.relerr(got, want) < 1e-13
    monkeypatch.setenv("TN_GRAM_F64_CP8", "1")          # 8-byte copies only: the same arithmetic on the same values
    assert np.array_equal(run(), got)
    monkeypatch.delenv("TN_GRAM_F64_CP8")
    monkeypatch.setenv("TN_GRAM_F64_UNFACTORED", "1")
    old = run()
    assert gu.relerr(old, want) < 1e-13
    assert gu.relerr(got, old) < 1e-14
