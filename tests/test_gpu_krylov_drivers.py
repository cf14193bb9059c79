"""The on-device Krylov drivers of libtn_b200.so (tn_cg / tn_minres / tn_lanczos, csrc/krylov.cu) and the exact refinement of
the tensor-core Gram modes, through the C ABI on a B200, against dense numpy / SciPy solves of the same systems and against
the CPU restatements of the recurrences (tests/fake_ops.py)."""
import numpy as np
import pytest
import torch

import fake_ops
import golden_util as gu

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)

from tensornetworksfork_b200 import ops  # noqa: E402
from tensornetworksfork_b200.ops import Factor  # noqa: E402

DEV = "cuda"
SHAPES = [(700, 3, 4, 3, 1), (4000, 6, 2, 6, 1), (1500, 5, 3, 4, 3), (9000, 12, 5, 1, 1), (3000, 24, 2, 24, 1)]


def make(rows, ma, mb, mc, V, seed, positive=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    Fa = torch.randn((rows * V, ma), generator=g)
    Fb = torch.rand((rows, mb), generator=g) * 2 - 1
    Fc = torch.randn((rows, mc), generator=g)
    w = torch.rand((rows * V,), generator=g) + 0.1 if positive else torch.randn((rows * V,), generator=g)
    return (Fa, Fb, Fc, w), rows * V, V


def factors(t, V, dev):
    Fa, Fb, Fc, w = (u.to(dev) for u in t)
    return (Factor(Fa, m=Fa.shape[1]), Factor(Fb, m=Fb.shape[1], div=V), Factor(Fc, m=Fc.shape[1], div=V)), w


def dense(t, V):
    Fa, Fb, Fc, w = (u.numpy() for u in t)
    rows = Fa.shape[0]
    idx = np.arange(rows) // V
    J = np.einsum("sa,sb,sc->sabc", Fa, Fb[idx], Fc[idx]).reshape(rows, -1)
    return (J * w[:, None]).T @ J


@pytest.mark.parametrize("shape", SHAPES)
def test_cg_with_and_without_preconditioner(shape):
    t, rows, V = make(*shape, seed=sum(shape))
    A0 = dense(t, V)
    P = A0.shape[0]
    sigma = np.abs(np.diag(A0)).mean()
    ridge = 0.5
    A = A0 / sigma + ridge * np.eye(P)
    rng = np.random.default_rng(1)
    b = rng.normal(size=P)
    want = np.linalg.solve(A, b)
    facs, w = factors(t, V, DEV)
    op = ops.Operator(P, factors=facs, w=w, rows=rows, sigma=torch.tensor([sigma], device=DEV), ridge=ridge)
    x, stats = ops.cg(op, torch.tensor(b, device=DEV), max_iter=4 * P, rtol=1e-12)
    rel, iters, stopped, applies = stats.tolist()[:4]
    assert stopped == 1.0 and rel <= 1e-12 and gu.relerr(x.cpu().numpy(), want) < 1e-9, (rel, iters)
    # preconditioned by the factor of a PERTURBED matrix (1e-4: coarser than the 3xTF32 Gram): same solution, few iterations
    Ap = A * (1.0 + 1e-4 * rng.normal(size=A.shape))
    Ap = 0.5 * (Ap + Ap.T)
    lda = (P + 7) // 8 * 8
    Ad = torch.zeros((P, lda), device=DEV)
    Ad[:, :P] = torch.tensor(Ap, device=DEV)
    work, info = ops.cholesky_factor(Ad)
    assert int(info.item()) == 0
    x2, stats2 = ops.cg(op, torch.tensor(b, device=DEV), precond=(Ad, work, info), max_iter=50, rtol=1e-12)
    rel2, iters2, crit2 = stats2[0].item(), stats2[1].item(), stats2[4].item()
    assert crit2 <= 1e-12 and iters2 <= 12 and gu.relerr(x2.cpu().numpy(), want) < 1e-10, (rel2, iters2, crit2)
    # the factor applied alone is the solve of the perturbed system
    y = ops.cholesky_apply(Ad, work, info, torch.tensor(b, device=DEV))
    assert gu.relerr(y.cpu().numpy(), np.linalg.solve(Ap, b)) < 1e-9


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_cg_matches_the_cpu_restatement_iteration_by_iteration(shape):
    """Same x0, same iteration cap, tolerance never reached: the device recurrence and the torch restatement walk the same path."""
    t, rows, V = make(*shape, seed=7 + sum(shape))
    P = t[0].shape[1] * t[1].shape[1] * t[2].shape[1]
    g = torch.Generator().manual_seed(3)
    b = torch.randn((P,), generator=g)
    x0 = torch.randn((P,), generator=g)
    facs_c, w_c = factors(t, V, "cpu")
    facs_d, w_d = factors(t, V, DEV)
    for iters in (1, 3, 7):
        op_c = ops.Operator(P, matvec=lambda v: fake_ops.matvec(*facs_c, w_c, rows, v), device="cpu")
        want, st_c = fake_ops.cg(op_c, b, x0=x0, max_iter=iters, rtol=0.0)
        op_d = ops.Operator(P, factors=facs_d, w=w_d, rows=rows)
        got, st_d = ops.cg(op_d, b.to(DEV), x0=x0.to(DEV), max_iter=iters, rtol=0.0)
        assert gu.relerr(got.cpu().numpy(), want.numpy()) < 1e-9
        assert abs(st_d[0].item() - st_c[0].item()) <= 1e-8 * max(1.0, st_c[0].item()) and int(st_d[1].item()) == iters


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_minres_matches_scipy_and_the_cpu_restatement(shape):
    from scipy.sparse.linalg import minres as sp_minres
    t, rows, V = make(*shape, seed=11 + sum(shape), positive=False)       # signed weights: symmetric indefinite operator
    A = dense(t, V)
    P = A.shape[0]
    rng = np.random.default_rng(2)
    b = rng.normal(size=P)
    facs_c, w_c = factors(t, V, "cpu")
    facs_d, w_d = factors(t, V, DEV)
    for iters in (2, 5, min(12, P // 3)):       # well below P: past that the Lanczos basis has lost orthogonality and iterates differ
        op_c = ops.Operator(P, matvec=lambda v: fake_ops.matvec(*facs_c, w_c, rows, v), device="cpu")
        want, st_c = fake_ops.minres(op_c, torch.tensor(b), max_iter=iters, rtol=1e-10)
        got, st_d = ops.minres(ops.Operator(P, factors=facs_d, w=w_d, rows=rows), torch.tensor(b, device=DEV), max_iter=iters, rtol=1e-10)
        sp, _ = sp_minres(A, b, maxiter=iters, rtol=1e-10)
        assert int(st_d[1].item()) == int(st_c[1].item())
        assert gu.relerr(got.cpu().numpy(), want.numpy()) < 1e-7, iters
        # SciPy's own recurrence (different stopping bookkeeping, same Krylov iterate)
        assert np.linalg.norm(got.cpu().numpy() - sp) <= 1e-6 * max(1.0, np.linalg.norm(sp)), iters


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_lanczos_matches_the_reference_recurrence(shape):
    t, rows, V = make(*shape, seed=13 + sum(shape))
    P = t[0].shape[1] * t[1].shape[1] * t[2].shape[1]
    g = torch.Generator().manual_seed(5)
    b = torch.randn((P,), generator=g)
    x0 = torch.randn((P,), generator=g)
    facs_c, w_c = factors(t, V, "cpu")
    facs_d, w_d = factors(t, V, DEV)
    for iters in (1, 2, 6, min(P, 12)):
        op_c = ops.Operator(P, matvec=lambda v: fake_ops.matvec(*facs_c, w_c, rows, v), device="cpu")
        want, st_c = fake_ops.lanczos(op_c, b, x0=x0, max_iter=iters, tol=1e-12)
        got, st_d = ops.lanczos(ops.Operator(P, factors=facs_d, w=w_d, rows=rows), b.to(DEV), x0=x0.to(DEV), max_iter=iters, tol=1e-12)
        assert int(st_d[1].item()) == int(st_c[1].item())
        assert gu.relerr(got.cpu().numpy(), want.numpy()) < 1e-7, (iters, gu.relerr(got.cpu().numpy(), want.numpy()))


def test_callback_operator_and_early_stop():
    """A Python matvec (what conv-TT / cum-sum hand over) driven by the same device recurrences; the tolerance stops the loop."""
    g = torch.Generator().manual_seed(0)
    P = 300
    Q = torch.randn((P, P), generator=g)
    A = (Q @ Q.t() / P + torch.eye(P)).to(DEV)
    b = torch.randn((P,), generator=g).to(DEV)
    calls = [0]

    def mv(v):
        calls[0] += 1
        return A @ v

    want = torch.linalg.solve(A, b)
    for fn in (ops.cg, ops.minres):
        calls[0] = 0
        x, stats = fn(ops.Operator(P, matvec=mv, device=b.device), b, max_iter=500, rtol=1e-10)
        rel, iters, stopped, applies = stats.tolist()[:4]
        assert stopped == 1.0 and iters < 100 and float((x - want).norm() / want.norm()) < 1e-8
        assert calls[0] == int(applies) and calls[0] <= iters + 8          # polled every few iterations, not run to max_iter
    x, stats = ops.lanczos(ops.Operator(P, matvec=mv, device=b.device), b, x0=torch.zeros_like(b), max_iter=60, tol=1e-9)
    assert float((x - want).norm() / want.norm()) < 1e-6

    def boom(v):
        raise RuntimeError("matvec failed")

    with pytest.raises(RuntimeError, match="matvec failed"):
        ops.cg(ops.Operator(P, matvec=boom, device=b.device), b, max_iter=5, rtol=1e-10)


def test_gram_trace_is_the_exact_trace():
    for shape in SHAPES:
        t, rows, V = make(*shape, seed=17 + sum(shape), positive=False)
        A = dense(t, V)
        facs, w = factors(t, V, DEV)
        tr = ops.gram_trace(*facs, w, rows).cpu().numpy()
        Aabs = dense((t[0], t[1], t[2], t[3].abs()), V)
        assert abs(tr[0] - np.trace(A)) <= 1e-11 * np.trace(Aabs) and abs(tr[1] - np.trace(Aabs)) <= 1e-11 * np.trace(Aabs)


@pytest.mark.parametrize("gram_mode", ["tf32x3", "tf32", "f16"])
@pytest.mark.parametrize("name", ["tt_poly_reg", "tnml_poly_xe", "cpd_reg", "tnml_sincos_qr", "tt_poly5_full"])
def test_tensor_core_modes_are_exact_after_refinement(name, gram_mode):
    """A free-running sweep in a tensor-core Gram mode against the REFERENCE recording at fp64 tolerances: the Gram (1e-5 /
    3e-4 accurate) only preconditions; losses, cores and predictions are those of the fp64 system."""
    import test_gpu_golden as tg
    fx = gu.load(name)
    meta = fx["meta"]
    layer = tg.build(fx)
    tg.set_cores(layer, fx["cores0"])
    tn = layer.tensor_network
    tn.gram_mode = gram_mode
    tn.small_site_fp64 = 0          # tiny fixtures: keep them on the refined path
    x, y = tg.data(fx)
    trace = []
    ok = tn.accumulating_swipe(x, y, tg.loss_of(meta), batch_size=meta["batch_size"], num_swipes=meta["num_swipes"], lr=meta["lr"],
                               method=meta["method"], eps=meta["eps"], eps_decay=meta.get("eps_decay"),
                               orthonormalize=meta.get("orthonormalize", False), skip_second=meta.get("skip_second", False),
                               loss_callback=lambda NS, node, l: trace.append((NS, tn.train_nodes.index(node), l)),
                               **gu.sweep_extras(meta))
    assert ok == fx["ok"]
    for (_, _, l), u in zip(trace, fx["updates"]):
        assert abs(l - u["loss"]) <= 1e-7 * max(1.0, abs(u["loss"])), (l, u["loss"])
    pred = tn.forward_batch(x, meta["batch_size"]).cpu().numpy()
    assert gu.relerr(pred.reshape(fx["pred"].shape), fx["pred"]) < 1e-7
    assert tn.solve_stats["refined"] + tn.solve_stats["gram_fp64_fallback"] == len(trace), tn.solve_stats
    assert tn.solve_stats["refined"] >= len(trace) - 1, tn.solve_stats


@pytest.mark.parametrize("case", [(5000, 24, 2, 24, 1, "sincos"), (3001, 38, 2, 38, 10, "sincos"), (700, 3, 4, 3, 1, None), (2000, 6, 2, 6, 1, None),
                                  (4000, 12, 3, 9, 3, None), (9000, 32, 4, 64, 1, "poly"), (300, 24, 2, 24, 1, "sincos"), (6000, 38, 6, 38, 1, "poly")])
def test_one_launch_matvec_matches_dense(case):
    """tn_matvec_kr3: the fused kernel (both passes on DMMA, factors read once; csrc/matvec_fused.cu) for small cores, the two-pass
    path otherwise (few rows, wide cores) -- against the dense J^T diag(w) J v, with feature maps and class rows (row divisors)."""
    rows, ma, mb, mc, V, fmap = case
    g = torch.Generator().manual_seed(sum(case[:5]))
    S = rows
    Fa = torch.randn((S * V, ma), generator=g)
    Fc = torch.randn((S, mc), generator=g)
    w = torch.randn((S * V,), generator=g)
    if fmap is None:
        Xb = torch.rand((S, mb), generator=g) * 2 - 1
        fb_c = Factor(Xb, m=mb, div=V)
        fb_d = Factor(Xb.to(DEV), m=mb, div=V)
    else:
        X = torch.rand((S, 7), generator=g) * 2 - 1
        kind = ops.MAP_SINCOS if fmap == "sincos" else ops.MAP_POLY
        fb_c = Factor(X, m=mb, div=V, map_kind=kind, col=3)
        fb_d = Factor(X.to(DEV), m=mb, div=V, map_kind=kind, col=3)
    v = torch.randn((ma * mb * mc,), generator=g)
    want = fake_ops.matvec(Factor(Fa, m=ma), fb_c, Factor(Fc, m=mc, div=V), w, S * V, v)
    got = ops.matvec(Factor(Fa.to(DEV), m=ma), fb_d, Factor(Fc.to(DEV), m=mc, div=V), w.to(DEV), S * V, v.to(DEV))
    torch.cuda.synchronize()
    assert gu.relerr(got.cpu().numpy(), want.numpy()) < 1e-12
    got2 = ops.matvec(Factor(Fa.to(DEV), m=ma), fb_d, Factor(Fc.to(DEV), m=mc, div=V), None, S * V, v.to(DEV))
    want2 = fake_ops.matvec(Factor(Fa, m=ma), fb_c, Factor(Fc, m=mc, div=V), None, S * V, v)
    assert gu.relerr(got2.cpu().numpy(), want2.numpy()) < 1e-12
