"""Mixed-precision local solve (tensor-core factorisation + fp64 refinement) against LAPACK on the same system, and the
large-system substitution kernels of the fp64 path.  Reference: TensorNetwork.solve_system, tensor/network.py:293-327."""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)

from tensornetworksfork_b200 import ops  # noqa: E402

DEV = "cuda"


def spd_device(P, seed, ridge=0.5, rank_extra=10):
    """A = B B^T / P + ridge I built on the device (fp64), returned padded (P x lda) plus a dense copy."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    B = torch.randn((P, P + rank_extra), device=DEV, generator=g)
    A = B @ B.t() / P
    A.diagonal().add_(ridge)
    lda = (P + 7) // 8 * 8
    Ap = torch.zeros((P, lda), device=DEV)
    Ap[:, :P] = A
    rhs = torch.randn((P,), device=DEV, generator=g)
    return A, Ap, rhs


@pytest.mark.parametrize("P", [1500, 3001, 5000, 9000])
def test_mixed_solve_matches_fp64(P):
    A, Ap, rhs = spd_device(P, P)
    want = torch.linalg.solve(A, rhs)
    x = rhs.clone()
    info, stats = ops.cholesky_solve_mixed(Ap, x, rtol=1e-12, max_iter=12)
    rel, iters = stats.tolist()
    assert int(info.item()) == 0
    assert rel <= 1e-12, (rel, iters)
    assert 1 <= iters <= 8, iters                       # the 3xTF32 factor is a strong preconditioner, not an exact one
    res = float(torch.norm(A @ x - rhs) / torch.norm(rhs))
    assert res < 1e-11, res
    assert float(torch.norm(x - want) / torch.norm(want)) < 1e-10
    # the strict upper triangle (the operator of the refinement) is untouched
    assert torch.equal(torch.triu(Ap[:, :P], 1), torch.triu(A, 1))


def test_mixed_factor_accuracy():
    """The tensor-core factor itself: L L^T reproduces A to the 3xTF32 level (fp32 accumulation over the panel width)."""
    P = 4000
    A, Ap, rhs = spd_device(P, 7)
    info, _ = ops.cholesky_solve_mixed(Ap, rhs.clone(), rtol=1e-6, max_iter=1)
    assert int(info.item()) == 0
    L = torch.tril(Ap[:, :P])
    err = float(torch.norm(L @ L.t() - A) / torch.norm(A))
    assert err < 5e-6, err
    assert err > 1e-12          # i.e. the tensor-core path really ran (an fp64 factor would be ~1e-16)


def test_mixed_reports_non_spd():
    P = 2000
    A, Ap, rhs = spd_device(P, 3)
    Ap[1500, 1500] = -1.0
    x = rhs.clone()
    info, stats = ops.cholesky_solve_mixed(Ap, x)
    assert int(info.item()) == 1501
    assert torch.equal(x, rhs)                          # no solution is written on failure


def test_mixed_ill_conditioned_still_converges_or_says_so():
    """cond ~ 1e7: the 1e-5 factor is a weak preconditioner; the call must either reach the residual or report it."""
    P = 2048
    g = torch.Generator(device=DEV).manual_seed(11)
    Q, _ = torch.linalg.qr(torch.randn((P, P), device=DEV, generator=g))
    lam = torch.logspace(0, -7, P, device=DEV)
    A = (Q * lam) @ Q.t()
    A = 0.5 * (A + A.t())
    lda = P
    Ap = A.clone()
    rhs = torch.randn((P,), device=DEV, generator=g)
    x = rhs.clone()
    info, stats = ops.cholesky_solve_mixed(Ap, x, rtol=1e-10, max_iter=30)
    rel, iters = stats.tolist()
    if int(info.item()) == 0 and rel <= 1e-10:
        assert float(torch.norm(A @ x - rhs) / torch.norm(rhs)) < 1e-9
    else:
        assert int(info.item()) != 0 or rel > 1e-10


@pytest.mark.parametrize("P", [8200, 9001])
def test_cholesky_large_substitution(P):
    """P > 8192 takes the 512-wide super-block substitution kernels."""
    A, Ap, rhs = spd_device(P, P + 1)
    x = rhs.clone()
    info = ops.cholesky_solve(Ap, x)
    assert int(info.item()) == 0
    res = float(torch.norm(A @ x - rhs) / torch.norm(rhs))
    assert res < 1e-11, res


@pytest.mark.parametrize("P", [8200, 9001, 12345])
def test_block_inverse_substitution_matches_serial_block_solve(P, monkeypatch):
    """The substitutions' 512 x 512 block solves as products with the inverted diagonal blocks (solve.cu::blkinv512_kernel,
    trsv_gemv512_kernel) against the serial block kernel (TN_TRSV_NO_BLKINV=1) and against the residual of the system; window
    counts that are not multiples of 512 / 64, repeated applications of one factor (the preconditioner of tn_cg)."""
    A, Ap, rhs = spd_device(P, P + 3)
    L = Ap.clone()
    work, info = ops.cholesky_factor(L, tensor_core=False)
    assert int(info.item()) == 0
    x = ops.cholesky_apply(L, work, info, rhs.clone())
    assert float(torch.norm(A @ x - rhs) / torch.norm(rhs)) < 1e-11
    x2 = ops.cholesky_apply(L, work, info, rhs.clone())
    assert float(torch.norm(x - x2) / torch.norm(x)) < 1e-13          # the backward panels combine row groups with fp64 atomics: equal to rounding
    monkeypatch.setenv("TN_TRSV_NO_BLKINV", "1")
    y = ops.cholesky_apply(L, work, info, rhs.clone())
    assert float(torch.norm(x - y) / torch.norm(y)) < 1e-12


def test_sweep_with_mixed_solve_tracks_fp64_solve():
    """A TT sweep whose local solves go through the mixed path lands on the same model as the fp64 solve."""
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(5)
    N, F = 6000, 11
    X = np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1)
    y = np.tanh(X[:, :1] + X[:, 1:2] * X[:, 2:3])
    preds = {}
    for mode in ("fp64", "mixed"):
        layer = tnb.TensorTrainLayer(3, 12, F + 1, output_shape=1, constrict_bond=False, perturb=True, seed=42)
        layer.to(DEV)
        net = layer.tensor_network
        net.solve_mode = mode
        xs, ys = torch.tensor(X, device=DEV), torch.tensor(y, device=DEV)
        ok = net.accumulating_swipe(xs, ys, tnb.SquareBregFunction(), batch_size=-1, num_swipes=1, method="ridge_cholesky",
                                    eps=1.0, eps_decay=0.5)
        assert ok
        preds[mode] = net.forward(xs, to_tensor=True).cpu().numpy()
        if mode == "mixed":
            assert net.solve_stats["mixed"] >= 1, net.solve_stats        # middle core: P = 12*12*12 = 1728
    assert gu.relerr(preds["mixed"], preds["fp64"]) < 1e-8


def test_mixed_solve_fallback_path_redoes_the_system_in_fp64():
    """Forcing the acceptance test to fail exercises the loud fall-back: A is re-expanded (its lower triangle was overwritten by the
    tensor-core factor) and solved by the fp64 Cholesky; the sweep must land exactly where the fp64 solve mode lands."""
    import tensornetworksfork_b200 as tnb
    rng = np.random.default_rng(6)
    N, F = 3000, 9
    X = np.concatenate([rng.uniform(-1, 1, size=(N, F)), np.ones((N, 1))], 1)
    y = np.tanh(X[:, :1] - X[:, 1:2] * X[:, 2:3])
    preds = {}
    for mode in ("fp64", "mixed"):
        layer = tnb.TensorTrainLayer(3, 10, F + 1, output_shape=1, constrict_bond=False, perturb=True, seed=1)
        layer.to(DEV)
        net = layer.tensor_network
        net.solve_mode = mode
        net.mixed_accept = -1.0            # nothing is ever accepted
        xs, ys = torch.tensor(X, device=DEV), torch.tensor(y, device=DEV)
        assert net.accumulating_swipe(xs, ys, tnb.SquareBregFunction(), batch_size=-1, num_swipes=1, method="ridge_cholesky", eps=[1.0, 0.1])
        preds[mode] = net.forward(xs, to_tensor=True).cpu().numpy()
        if mode == "mixed":
            assert net.solve_stats["mixed_fallback"] >= 1 and net.solve_stats["mixed"] == 0, net.solve_stats
            assert net._mixed_floor == 2.0          # the largest ridge (2 eps) that failed is remembered ...
    assert np.array_equal(preds["mixed"], preds["fp64"])       # ... and the redo is the plain fp64 path, bit for bit


def test_kernels_accept_empty_batches():
    """rows == 0 (an empty shard / empty last batch) is a no-op, not an error."""
    from tensornetworksfork_b200.ops import Factor
    e = torch.empty((0, 4), device=DEV)
    x = torch.empty((0, 3), device=DEV)
    core = torch.randn((4, 3, 5), device=DEV)
    assert tuple(ops.env_update(e, Factor(x, m=3), core, 0).shape) == (0, 5)
    fa, fb, fc = Factor(e, m=4), Factor(x, m=3), Factor(torch.empty((0, 2), device=DEV), m=2)
    w = torch.empty((0,), device=DEV)
    for mode in (ops.GRAM_FP64, ops.GRAM_TF32X3):
        M = ops.gram(mode, fa, fb, fc, w, 0)
        assert M.numel() == 10 * 6 * 3 and float(M.abs().sum()) == 0.0
    b = ops.rhs(fa, fb, fc, w, 0)
    assert b.numel() == 24 and float(b.abs().sum()) == 0.0
