"""tcgen05 Gram kernel (TF32 / 3xTF32, TMEM accumulators) against the fp64 kernel and numpy (needs a B200).

Stated tolerances (relative Frobenius error of the unique-entry tensor M against fp64):
  3xTF32 : <= 3e-5   products are fp32-grade (~3e-7), but tcgen05 accumulates in fp32 with TRUNCATION (measured:
                     error grows linearly with the rows accumulated between two fp64 flushes, ~7e-9 per row; 1024
                     rows per flush by default; 2048 in the library -> ~1.2e-5).  The end-to-end bound that matters -- per-update loss of a
                     free-running sweep within 1e-6 of the fp64 mode -- is asserted at the bottom of this file.
  TF32   : <= 2e-3
"""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)

from tensornetworksfork_b200 import ops  # noqa: E402
from tensornetworksfork_b200.ops import Factor  # noqa: E402

DEV = "cuda"
TOL = {ops.GRAM_TF32X3: 3e-5, ops.GRAM_TF32: 2e-3, ops.GRAM_F16: 5e-3}      # fp16 operands: every product rounded to 11 bits

SHAPES = [
    # rows, ma, mb, mc, V
    (64, 2, 2, 2, 1),          # one tile, one chunk window
    (1000, 4, 5, 4, 1),        # ragged rows
    (5000, 6, 9, 6, 1),        # several U tiles, flush boundary (4096) crossed
    (20000, 24, 2, 24, 1),     # config-3 middle site: BN=256, two V tiles, T=2
    (3000, 38, 6, 38, 1),      # config-5b middle site: 3 V tiles
    (2500, 1, 29, 38, 1),      # config-5a first site (trivial left factor)
    (2500, 38, 29, 1, 1),      # config-5a last site
    (9000, 100, 9, 1, 1),      # config-2 CPD factor: rank 100
    (1200, 5, 3, 4, 3),        # class rows, signed weights
]


def make(rows, ma, mb, mc, V, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    S = rows
    Fa = torch.randn((S * V, ma), device=DEV, generator=g)
    Fb = torch.rand((S, mb), device=DEV, generator=g) * 2 - 1
    Fc = torch.randn((S, mc), device=DEV, generator=g)
    w = torch.randn((S * V,), device=DEV, generator=g)
    return Factor(Fa, m=ma), Factor(Fb, m=mb, div=V), Factor(Fc, m=mc, div=V), w, S * V


@pytest.mark.parametrize("mode", [ops.GRAM_TF32X3, ops.GRAM_TF32, ops.GRAM_F16])
@pytest.mark.parametrize("shape", SHAPES)
def test_tc_gram_matches_fp64(shape, mode):
    fa, fb, fc, w, rows = make(*shape, seed=sum(shape))
    ref = ops.gram(ops.GRAM_FP64, fa, fb, fc, w, rows)
    got = ops.gram(mode, fa, fb, fc, w, rows)
    torch.cuda.synchronize()
    err = gu.relerr(got.cpu().numpy(), ref.cpu().numpy())
    assert err < TOL[mode], err


def test_tc_gram_positive_weights_and_accumulate():
    fa, fb, fc, w, rows = make(6000, 8, 4, 8, 1, seed=5)
    w = w.abs()
    ref = ops.gram(ops.GRAM_FP64, fa, fb, fc, w, rows)
    got = ops.gram(ops.GRAM_TF32X3, fa, fb, fc, w, rows)
    assert gu.relerr(got.cpu().numpy(), ref.cpu().numpy()) < 3e-5
    twice = ops.gram(ops.GRAM_TF32X3, fa, fb, fc, w, rows, M=got.clone(), accumulate=True)
    assert gu.relerr(twice.cpu().numpy(), 2 * ref.cpu().numpy()) < 3e-5


def test_tc_gram_feature_map_factor():
    g = torch.Generator(device=DEV).manual_seed(9)
    S = 7000
    X = torch.rand((S, 12), device=DEV, generator=g) * 2 - 1
    L = torch.randn((S, 16), device=DEV, generator=g)
    R = torch.randn((S, 16), device=DEV, generator=g)
    w = torch.full((S,), 2.0, device=DEV)
    fb = Factor(X, m=2, map_kind=ops.MAP_SINCOS, col=7)
    ref = ops.gram(ops.GRAM_FP64, Factor(L, m=16), fb, Factor(R, m=16), w, S)
    got = ops.gram(ops.GRAM_TF32X3, Factor(L, m=16), fb, Factor(R, m=16), w, S)
    assert gu.relerr(got.cpu().numpy(), ref.cpu().numpy()) < 3e-5


def test_sweep_in_3xtf32_tracks_fp64_sweep():
    """Free-running sweep (well-conditioned, eps >= 0.25): per-update loss and predictions of the 3xTF32 Gram
    mode against the fp64 mode, tolerance 1e-6 relative (BASELINE.json north_star)."""
    import tensornetworksfork_b200 as tnb
    g = torch.Generator(device=DEV).manual_seed(3)
    N, F = 20000, 6
    X = torch.cat([torch.rand((N, F), device=DEV, generator=g) * 2 - 1, torch.ones((N, 1), device=DEV)], 1)
    y = torch.tanh(X[:, :1]) + 0.5 * X[:, 1:2] * X[:, 2:3] + 0.05 * torch.randn((N, 1), device=DEV, generator=g)
    out = {}
    for mode in ("fp64", "tf32x3"):
        layer = tnb.TensorTrainLayer(4, 8, F + 1, output_shape=1, constrict_bond=False, seed=42).to(DEV)
        layer.tensor_network.gram_mode = mode
        losses = []
        ok = layer.tensor_network.accumulating_swipe(X, y, tnb.SquareBregFunction(), num_swipes=2, method="ridge_cholesky", eps=1.0,
                                                     eps_decay=0.5, loss_callback=lambda NS, n, l: losses.append(l))
        assert ok
        out[mode] = (losses, layer.tensor_network.forward(X, to_tensor=True))
    for a, b in zip(out["tf32x3"][0], out["fp64"][0]):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (a, b)
    assert gu.relerr(out["tf32x3"][1].cpu().numpy(), out["fp64"][1].cpu().numpy()) < 1e-5


def test_tc_gram_wide_dynamic_range():
    """Factors of long normalised chains are tiny (1e-30) or large; the fp32 staging is range-scaled by powers of two."""
    fa, fb, fc, w, rows = make(4000, 12, 3, 12, 1, seed=77)
    fa = Factor(fa.tensor * 1e-28, m=12)
    fc = Factor(fc.tensor * 3e17, m=12)
    w = w.abs() * 1e-9
    ref = ops.gram(ops.GRAM_FP64, fa, fb, fc, w, rows)
    got = ops.gram(ops.GRAM_TF32X3, fa, fb, fc, w, rows)
    assert torch.isfinite(got).all()
    assert gu.relerr(got.cpu().numpy(), ref.cpu().numpy()) < 3e-5


def test_tc_gram_config5a_site_full_width():
    """BASELINE config-5a middle site (r=38, f=29: 741 x 435 x 741 unique entries, 1.9 GB) at 32 768 rows: the 3xTF32 tensor-core
    Gram against the fp64 kernel, and additivity over row shards (what the multi-GPU all-reduce relies on)."""
    S, ma, mb, mc = 32768, 38, 29, 38
    g = torch.Generator(device=DEV).manual_seed(5)
    Fa = torch.randn((S, ma), device=DEV, generator=g)
    Fb = torch.rand((S, mb), device=DEV, generator=g) * 2 - 1
    Fc = torch.randn((S, mc), device=DEV, generator=g)
    w = torch.full((S,), 2.0, device=DEV)
    fac = lambda lo, hi: (Factor(Fa[lo:hi], m=ma), Factor(Fb[lo:hi], m=mb), Factor(Fc[lo:hi], m=mc))
    full = ops.gram(ops.GRAM_TF32X3, *fac(0, S), w, S)
    ref = ops.gram(ops.GRAM_FP64, *fac(0, S), w, S)
    err = float((full - ref).norm() / ref.norm())
    assert err < 3e-5, err
    half = S // 2 + 8
    parts = ops.gram(ops.GRAM_TF32X3, *fac(0, half), w[:half], half)
    parts = ops.gram(ops.GRAM_TF32X3, *fac(half, S), w[half:], S - half, M=parts, accumulate=True)
    assert float((parts - ref).norm() / ref.norm()) < 3e-5
    # the expanded system is symmetric positive semi-definite by construction: check a random Rayleigh quotient via matvec
    v = torch.randn(ma * mb * mc, device=DEV, generator=g)
    Av = ops.matvec(*fac(0, 4096), w[:4096], 4096, v)
    assert float(torch.dot(v, Av)) >= 0.0


def test_f16_gram_special_cases():
    """FP16-operand mode (kind::f16, mode 3): feature-map factor, class rows with signed weights, accumulation over row shards, a
    wide dynamic range (per-factor power-of-two scaling keeps it finite) and the config-5a middle site."""
    g = torch.Generator(device=DEV).manual_seed(9)
    S = 7000
    X = torch.rand((S, 12), device=DEV, generator=g) * 2 - 1
    L = torch.randn((S, 16), device=DEV, generator=g)
    R = torch.randn((S, 16), device=DEV, generator=g)
    w = torch.full((S,), 2.0, device=DEV)
    fb = Factor(X, m=2, map_kind=ops.MAP_SINCOS, col=7)
    ref = ops.gram(ops.GRAM_FP64, Factor(L, m=16), fb, Factor(R, m=16), w, S)
    got = ops.gram(ops.GRAM_F16, Factor(L, m=16), fb, Factor(R, m=16), w, S)
    assert gu.relerr(got.cpu().numpy(), ref.cpu().numpy()) < 5e-3
    half = S // 2 + 8
    parts = ops.gram(ops.GRAM_F16, Factor(L[:half], m=16), Factor(X[:half], m=2, map_kind=ops.MAP_SINCOS, col=7), Factor(R[:half], m=16), w[:half], half)
    parts = ops.gram(ops.GRAM_F16, Factor(L[half:], m=16), Factor(X[half:], m=2, map_kind=ops.MAP_SINCOS, col=7), Factor(R[half:], m=16), w[half:],
                     S - half, M=parts, accumulate=True)
    assert gu.relerr(parts.cpu().numpy(), ref.cpu().numpy()) < 5e-3
    fa, fb2, fc, w2, rows = make(4000, 12, 3, 12, 1, seed=77)
    fa = Factor(fa.tensor * 1e-28, m=12)
    fc = Factor(fc.tensor * 3e17, m=12)
    w2 = w2.abs() * 1e-9
    ref = ops.gram(ops.GRAM_FP64, fa, fb2, fc, w2, rows)
    got = ops.gram(ops.GRAM_F16, fa, fb2, fc, w2, rows)
    assert torch.isfinite(got).all() and gu.relerr(got.cpu().numpy(), ref.cpu().numpy()) < 5e-3
    S, ma, mb, mc = 16384, 38, 29, 38
    Fa = torch.randn((S, ma), device=DEV, generator=g)
    Fb = torch.rand((S, mb), device=DEV, generator=g) * 2 - 1
    Fc = torch.randn((S, mc), device=DEV, generator=g)
    w = torch.full((S,), 2.0, device=DEV)
    full = ops.gram(ops.GRAM_F16, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
    ref = ops.gram(ops.GRAM_FP64, Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
    assert float((full - ref).norm() / ref.norm()) < 5e-3


@pytest.mark.parametrize("shape", [(4096 + 70, 38, 29, 38), (6000, 24, 2, 24), (3000, 38, 6, 38), (2500, 38, 29, 1), (9000, 100, 9, 1),
                                   (1500, 7, 3, 5), (2000, 3, 40, 9)])
def test_f16_run_ordered_producers(shape, monkeypatch):
    """fp16 Gram kernel with run-ordered producers (gram_tc.cu::gram_tc16_run_kernel: a thread owns one 8-sample piece of eight
    consecutive tile rows and reuses the run prefix w fa[ia] fa[ja] fb[ib] / fc[ic]) against fp64 and against the row-per-thread
    kernel: same operands up to the order of the fp16 multiplications.  Shapes: the config-5a / 3 / 5b middle sites, a last site
    (one V column), the CPD factor, short runs (mB = 2, 3: more than two prefixes per eight rows -> the reloading path), long runs."""
    S, ma, mb, mc = shape
    g = torch.Generator(device=DEV).manual_seed(S + ma)
    Fa = torch.randn((S, ma), device=DEV, generator=g)
    Fb = torch.rand((S, mb), device=DEV, generator=g) * 2 - 1
    Fc = torch.randn((S, mc), device=DEV, generator=g)
    w = torch.rand((S,), device=DEV, generator=g) + 0.5
    args = (Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w, S)
    monkeypatch.setenv("TN_TC16_VIMG", "0")
    monkeypatch.setenv("TN_TC16_RUN", "0")
    one = ops.gram(ops.GRAM_F16, *args, flush_rows=2048)
    ref = ops.gram(ops.GRAM_FP64, *args)
    monkeypatch.setenv("TN_TC16_RUN", "1")
    two = ops.gram(ops.GRAM_F16, *args, flush_rows=2048)
    torch.cuda.synchronize()
    assert gu.relerr(two.cpu().numpy(), ref.cpu().numpy()) < 5e-3
    assert gu.relerr(two.cpu().numpy(), one.cpu().numpy()) < 2e-3
    again = ops.gram(ops.GRAM_F16, *args, flush_rows=2048)
    assert torch.equal(two, again)          # deterministic: every tile of M has one owner
    # pre-synthesised V images + pair products of the left factor (gram_tc16_vimg_kernel): operands rounded once from fp32 products
    monkeypatch.setenv("TN_TC16_VIMG", "1")
    three = ops.gram(ops.GRAM_F16, *args, flush_rows=2048)
    torch.cuda.synchronize()
    assert gu.relerr(three.cpu().numpy(), ref.cpu().numpy()) < 5e-3
    assert gu.relerr(three.cpu().numpy(), one.cpu().numpy()) < 2e-3
    assert torch.equal(three, ops.gram(ops.GRAM_F16, *args, flush_rows=2048))
    acc = ops.gram(ops.GRAM_F16, *args, flush_rows=2048, M=three.clone(), accumulate=True)
    assert gu.relerr(acc.cpu().numpy(), 2 * three.cpu().numpy()) < 1e-12
    # operand tiles in the 128-byte-swizzled K-major layout: the same operands in another arrangement -> the same bits
    monkeypatch.setenv("TN_TC16_SW128", "1")
    four = ops.gram(ops.GRAM_F16, *args, flush_rows=2048)
    torch.cuda.synchronize()
    assert torch.equal(three, four)
