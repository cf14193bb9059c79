import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
torch.set_default_dtype(torch.float64)
import conv_cases as cc, linear_cases as lc, krylov_cases as kc, growing_case as gc
for name in ["conv_lanczos_xe", "conv_lanczos_reg"]:
    for chunk in (None, 37):
        print(name, chunk, cc.run_case(name, "cuda", chunk_rows=chunk))
for name in ["conv_scipy_cg", "conv_scipy_minres", "conv_scipy_cg_2col"]:
    print(name, cc.run_case(name, "cuda", scipy_object=True), cc.run_case(name, "cuda", scipy_object=False, loss_prefix=8))
for name in sorted(lc.CASES):
    print(name, lc.run_case(name, "cuda"))
for name in ["krylov_lanczos_reg", "krylov_lanczos_xe"]:
    print(name, kc.run_case(name, "cuda"))
for tag in "ab":
    print("growing", tag, gc.run(tag, "cuda"))
# flows added after the round-1 GPU minutes were spent (tests/test_zz_gpu_late.py): measure, then tighten those tolerances
import batch_case as bc, gradient_case as grc
print("conv grow", cc.run_grow("cuda"))
for name in ["conv_type1", "conv_onecol", "conv_nocb", "conv_dense_xe", "conv_dense_reg"]:
    print(name, cc.run_case(name, "cuda"))
for tag in bc.CASES:
    print("batch", tag, bc.run(tag, "cuda"))
for name in grc.CASES:
    print(name, grc.run(name, "cuda"))
print("krylov_cumsum_lanczos", kc.run_case("krylov_cumsum_lanczos", "cuda"))
print("krylov_cumsum_cg", kc.run_case("krylov_cumsum_cg", "cuda", scipy_object=True))
# recordings at the BASELINE configurations' own sizes (tests/golden/make_golden_cfg*.py)
import cfg1_case, cfg2_case, cfg3_case, cfg4a_case, cfg4b_case, cfg5a_case, cfg5b_case
for gm in ("fp64", "tf32x3"):
    le, pe, ce = cfg1_case.run("cuda", gram_mode=gm); print("cfg1", gm, le, pe, ce)
    le, pe = cfg2_case.run("cuda", gram_mode=gm); print("cfg2", gm, le.max(), pe)
    le, pe = cfg5b_case.run("cuda", gram_mode=gm); print("cfg5b", gm, le.max(), pe)
    print("cfg5a dense", gm, cfg5a_case.dense("cuda", gm))
le, pe = cfg3_case.run("cuda"); print("cfg3", le.max(), pe)
print("cfg4a", [float(v.max()) if hasattr(v, "max") else v for v in cfg4a_case.run("cuda")])
print("cfg4b", [float(v.max()) if hasattr(v, "max") else v for v in cfg4b_case.run("cuda")])
print("cfg5a matrix-free", cfg5a_case.matrix_free("cuda"))
