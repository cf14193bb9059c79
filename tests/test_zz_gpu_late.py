"""GPU parity tests of the host flows added late in round 1 (needs a B200).

They drive the same kernels, at the same kinds of shapes, as the files before them -- only the host flow is new -- and each has a
twin on the CPU stand-in kernels (named in its docstring).  History: in round 1 they ran once behind non-strict markers and
``test_conv_growing_flow_gpu`` failed (after ``grow_cart`` the old last cores are stride-0 broadcasts, and a pixel-core Jacobian
handed such a core to ``tn_rows_dot``, whose wrapper asserts unit row stride); the CPU stand-ins now assert the kernels' layout
contracts (``tests/fake_ops.py``), ``conv._to_canon`` densifies broadcast cores, and since round 2 every test of this file runs
WITHOUT markers and passes on hardware (``profiles/r2_pytest_gpu_call24.log``), including the tensor-core twins of the BASELINE
configurations, which follow the reference recordings through the exact refinement.  The file sorts last so that it cannot hide
other results under ``pytest -x``.
"""
import os

import pytest
import torch

import conv_cases as cc

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def test_conv_growing_flow_gpu():
    """grow_cart between dense sweeps (image_convolution_growing_MNIST.py:84-103) against the reference recording
    tests/golden/conv_grow.npz; CPU twin: test_conv_cpu.py::test_conv_growing_flow_host_logic."""
    for pi, (fwd, core, loss, pred) in enumerate(cc.run_grow("cuda")):
        # after a growth the prediction starts from free-running cores: same bound as the final prediction of a sweep
        assert fwd < (1e-12 if pi == 0 else 1e-7) and core < 1e-7 and loss < 1e-9 and pred < 1e-7, (pi, fwd, core, loss, pred)


@pytest.mark.parametrize("tag", ["unique", "same", "block"])
def test_minibatch_estimator_gpu(tag):
    """TensorTrainBatchRegressor (reference tensor/module.py:308-500) against tests/golden/batch_tt.npz; CPU twin:
    test_growing_cpu.py::test_minibatch_estimator_swipe_methods (errors there: 1e-16 trajectory, 3e-14 cores)."""
    import batch_case as bc
    traj_err, pred_err, core_err = bc.run(tag, "cuda")
    assert traj_err < 1e-8 and pred_err < 1e-7 and core_err < 1e-6, (traj_err, pred_err, core_err)


@pytest.mark.parametrize("name", ["grad_tt_reg", "grad_tt_xe", "grad_type1"])
def test_gradient_method_gpu(name):
    """accumulating_swipe(method='gradient') (reference network.py:458-470, :558-584) against tests/golden/grad_*.npz: row-range
    right-hand sides per minibatch; CPU twin: test_host_sweep_cpu.py::test_gradient_method_matches_reference_recording."""
    import gradient_case as gcase
    core_err, loss_err = gcase.run(name, "cuda")
    assert core_err < 1e-10 and loss_err < 1e-10, (core_err, loss_err)


@pytest.mark.parametrize("name", ["conv_type1", "conv_onecol", "conv_nocb"])
def test_conv_type1_and_degenerate_columns_gpu(name):
    """Type-I image model (sum of conv-TTs with 1..3 columns, AAMNST.py:157-203), a single column, convolution_bond = -1, against
    their reference recordings; CPU twin: test_conv_cpu.py::test_conv_type1_and_degenerate_columns_host_logic."""
    fwd, core, loss, pred = cc.run_case(name, "cuda")
    assert fwd < 1e-12 and core < 1e-7 and loss < 1e-9 and pred < 1e-7, (fwd, core, loss, pred)


def test_cumsum_matrix_free_sweeps_gpu():
    """lanczos_swipe / scipy_swipe on the cum-sum train against tests/golden/krylov_cumsum_*.npz; CPU twins in test_krylov_cpu.py."""
    import krylov_cases as kc
    core_err, loss_err = kc.run_case("krylov_cumsum_lanczos", "cuda")
    # six Lanczos steps without re-orthogonalisation amplify rounding: the CPU twin sits at 8e-9 (cores) / 5e-10 (losses)
    assert core_err < 1e-6 and loss_err < 1e-7, (core_err, loss_err)
    core_err, loss_err = kc.run_case("krylov_cumsum_cg", "cuda", scipy_object=True)
    assert core_err < 5e-4 and loss_err < 5e-5, (core_err, loss_err)


@pytest.mark.parametrize("gram_mode", ["fp64", "tf32x3", "tf32", "f16"])
def test_baseline_config1_full_size_gpu(gram_mode):
    """BASELINE config 1 at full size against tests/golden/cfg1_full.npz (recorded from the unmodified reference): the first two
    half-sweeps (ridge 0.075, 2.8e-3) to rounding / to the 3xTF32 bound, the third (1e-4) loosely; beyond that the reference's own
    trajectory is chaotic and only the level of the final loss is compared.  CPU twin: test_host_sweep_cpu.py."""
    import cfg1_case as c1
    loss_err, pred_err, core_err = c1.run("cuda", gram_mode=gram_mode)
    # both modes: in 'tf32x3' the Gram only preconditions the exact refinement (network.py::_solve_refined), and the ridges it is too
    # coarse for (below ~1e-4 here) are redone with the fp64 Gram
    assert loss_err[:5].max() < 1e-9 and loss_err[5:7].max() < 1e-3, loss_err
    assert loss_err[-1] < 0.2, loss_err


@pytest.mark.parametrize("gram_mode", ["fp64", "tf32x3", "tf32", "f16"])
def test_baseline_config2_full_size_gpu(gram_mode):
    """BASELINE config 2 at full size (CPD rank 100, 20640 x 9, 5 factors, two sweeps with the wrapper's ridge schedule 1.0 * 0.5^NS)
    against tests/golden/cfg2_full.npz, recorded from the unmodified reference: all 17 per-update losses and the final prediction."""
    import cfg2_case as c2
    loss_err, pred_err = c2.run("cuda", gram_mode=gram_mode)
    tol = 1e-7          # both modes: the tensor-core Gram only preconditions the exact refinement
    assert loss_err.max() < tol and pred_err < 10 * tol, (loss_err, pred_err)


def test_baseline_config3_chain_gpu():
    """BASELINE config 3's 90-site chain (sin-cos map, rank 24, QR re-gauge) on a 4096-row subsample, one full sweep = 179 updates,
    against tests/golden/cfg3_chain90.npz recorded from the unmodified reference; fused feature map (what TNMLRegressor passes)."""
    import cfg3_case as c3
    loss_err, pred_err = c3.run("cuda")
    assert loss_err.max() < 1e-6 and pred_err < 1e-5, (loss_err.max(), pred_err)


@pytest.mark.parametrize("gram_mode", ["fp64", "tf32x3", "tf32", "f16"])
def test_baseline_config5b_chain_gpu(gram_mode):
    """BASELINE config 5b's chain (28 sites, polynomial degree 5, rank 38, QR re-gauge, P up to 8664 -- with 'tf32x3' the mixed
    tensor-core solve is on the path) on a 2048-row subsample, one sweep = 55 updates, against tests/golden/cfg5b_chain28.npz
    recorded from the unmodified reference."""
    import cfg5b_case as c5
    loss_err, pred_err = c5.run("cuda", gram_mode=gram_mode)
    tol = 1e-6          # both modes (north_star: <= 1e-6 for fp64 / 3xTF32)
    assert loss_err.max() < tol and pred_err < 10 * tol, (loss_err.max(), pred_err)


def test_baseline_config4b_full_model_size_gpu():
    """BASELINE config 4b at full model size (r = 38, CB = 4, 72 200-parameter patch core) under scipy_swipe(minres) against
    tests/golden/cfg4b_shape.npz recorded from the unmodified reference; CPU twin in test_conv_cpu.py (exact losses, 7e-16)."""
    import cfg4b_case as c4
    loss_err, pred_err = c4.run("cuda")
    assert loss_err.max() < 1e-5 and pred_err < 1e-4, (loss_err, pred_err)        # float32 Krylov recurrences on the host


@pytest.mark.parametrize("gram_mode", ["fp64", "tf32x3"])
def test_baseline_config5a_gram_fingerprint_gpu(gram_mode):
    """The north-star shape (5 cores, rank 38, 28 features + bias): the dense 41 876 x 41 876 system of the middle core, expanded from
    the unique entries the Gram kernel accumulates (fp64 DMMA / tcgen05 3xTF32), against b, diag(A), A v and the Frobenius norm of the
    reference's own get_A_b (tests/golden/cfg5a_gram.npz).  14 GB for A."""
    import cfg5a_case as c5
    b_err, diag_err, av_err, fro_err, asym = c5.dense("cuda", gram_mode)
    tol = 1e-12 if gram_mode == "fp64" else 3e-5
    assert b_err < 1e-12 and diag_err < tol and av_err < tol and fro_err < tol and asym < 1e-16, (b_err, diag_err, av_err, fro_err, asym)
    mf = c5.matrix_free("cuda")
    assert max(mf) < 1e-12, mf


def test_baseline_config4a_local_size_chain_gpu():
    """BASELINE config 4a at its full local size (rank 38, 9 logits, P up to 2888) on a 16-site chain under scipy_swipe(cg) against
    tests/golden/cfg4a_chain16.npz recorded from the unmodified reference; CPU twin in test_krylov_cpu.py (2e-16 / 7e-16)."""
    import cfg4a_case as c4
    loss_err, pred_err = c4.run("cuda")
    assert loss_err.max() < 1e-5 and pred_err < 1e-4, (loss_err.max(), pred_err)      # float32 Krylov recurrences on the host


@pytest.mark.parametrize("shape", [(64, 2, 2, 2, 1), (5000, 6, 9, 6, 1), (20000, 24, 2, 24, 1), (3000, 38, 6, 38, 1), (2500, 38, 29, 1, 1),
                                   (1200, 5, 3, 4, 3)])
def test_tc_gram_planar_raw_slots_match_the_row_major_layout(shape, monkeypatch):
    """The planar raw-factor ring filled by bulk copies (gram_tc.cu, the default since round 2: 588 vs 575 TF/s on the config-5a
    middle site) must give the same M as the row-major cp.async ring (TN_TC_RAW_ROWMAJOR=1) -- only shared-memory addresses change,
    not the arithmetic or its order -- or, where the factors are too small for the planar slot, fall back to it."""
    from tensornetworksfork_b200 import ops
    import test_gpu_gram_tc as tg
    fa, fb, fc, w, rows = tg.make(*shape, seed=sum(shape))
    monkeypatch.setenv("TN_TC_RAW_ROWMAJOR", "1")
    ref = ops.gram(ops.GRAM_TF32X3, fa, fb, fc, w, rows)
    monkeypatch.delenv("TN_TC_RAW_ROWMAJOR", raising=False)
    got = ops.gram(ops.GRAM_TF32X3, fa, fb, fc, w, rows)
    torch.cuda.synchronize()
    # fp64 atomics of different CTAs land in a different order from run to run: compare to rounding, not bit for bit
    assert float((got - ref).norm() / ref.norm()) < 1e-12


def test_qr_regauge_of_wide_cores_shrinks_the_bond_gpu():
    """An unconstricted 4-site train with r = 5 > f = 2: the first cores are wide (2 x 5, 4 x 5), the reference's reduced QR shrinks
    their bonds (network.py:644-657, 686-704).  The engine builds Q from the square leading block (tn_qr) and R = [R1 | Q^T A2]
    (tn_env_update); cores against the numpy oracle after a left and a right sweep of re-gauges, predictions unchanged
    (first run on a B200: profiles/r2_qr_wide_check.log; CPU twin: test_host_logic.py::test_qr_regauge_of_wide_cores_cpu)."""
    import qr_wide_case as qw
    mid, final, worst, drift = qw.run("cuda")
    assert mid[0] == (1, 2, 2) and mid[1] == (2, 2, 4), mid
    assert worst < 1e-11 and drift < 1e-11, (worst, drift)
