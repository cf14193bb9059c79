"""Characterise the tcgen05 Gram kernel: error vs flush window, and time on config-sized sites."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
DEV = "cuda"

def rel(a, b):
    return float((a - b).norm() / b.norm())

def make(S, ma, mb, mc, seed, wkind):
    g = torch.Generator(device=DEV).manual_seed(seed)
    Fa = torch.randn((S, ma), device=DEV, generator=g)
    Fb = torch.rand((S, mb), device=DEV, generator=g) * 2 - 1
    Fc = torch.randn((S, mc), device=DEV, generator=g)
    w = torch.randn((S,), device=DEV, generator=g) if wkind == "rand" else torch.full((S,), 2.0, device=DEV)
    return Factor(Fa, m=ma), Factor(Fb, m=mb), Factor(Fc, m=mc), w

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("err", "all"):
    for wkind in ("rand", "const"):
        fa, fb, fc, w = make(32768, 24, 2, 24, 1, wkind)
        ref = ops.gram(ops.GRAM_FP64, fa, fb, fc, w, 32768)
        for fr in (256, 512, 1024, 2048, 4096, 16384):
            os.environ["TN_TC_FLUSH_ROWS"] = str(fr)
            for mode, nm in ((ops.GRAM_TF32X3, "tf32x3"), (ops.GRAM_TF32, "tf32")):
                got = ops.gram(mode, fa, fb, fc, w, 32768)
                # diagonal-scaled error: relative to sqrt(M_ii M_jj) is what matters for the solve; report Frobenius
                print(json.dumps({"probe": "err", "w": wkind, "flush_rows": fr, "mode": nm, "rel_fro": rel(got, ref)}))
    os.environ.pop("TN_TC_FLUSH_ROWS", None)
if what in ("time", "all"):
    for name, S, ma, mb, mc in (("cfg5a_mid", 65536, 38, 29, 38), ("cfg5a_first", 262144, 1, 29, 38), ("cfg3_mid", 515345, 24, 2, 24),
                                ("cfg5b_mid", 262144, 38, 6, 38), ("cfg2", 20640, 100, 9, 1)):
        fa, fb, fc, w = make(S, ma, mb, mc, 2, "const")
        np_ = lambda m: m * (m + 1) // 2
        flops = 2.0 * S * np_(ma) * np_(mb) * np_(mc)
        for mode, nm, mult in ((ops.GRAM_TF32X3, "tf32x3", 3), (ops.GRAM_TF32, "tf32", 1)):
            for fr in (4096, 1024):
                os.environ["TN_TC_FLUSH_ROWS"] = str(fr)
                M = torch.empty(np_(ma) * np_(mb) * np_(mc), device=DEV)
                ms = timeit(lambda: ops.gram(mode, fa, fb, fc, w, S, M=M), n=2)
                print(json.dumps({"probe": "time", "site": name, "rows": S, "mode": nm, "flush_rows": fr, "ms": ms,
                                  "issued_tflops": flops * mult / ms / 1e9, "useful_tflops": flops / ms / 1e9}))
if what in ("chol",):
    for P in (2888, 8664, 20000, 41876):
        lda = (P + 7) // 8 * 8
        g = torch.Generator(device=DEV).manual_seed(P)
        A = torch.empty((P, lda), device=DEV)
        blk = 4096
        for i in range(0, P, blk):      # symmetric, diagonally dominant: cheap to generate at 14 GB
            A[i:i + blk, :P] = 0.5 / P ** 0.5 * torch.randn((min(blk, P - i), P), device=DEV, generator=g)
        A[:, :P] = 0.5 * (A[:, :P] + A[:, :P].t())
        A[:, :P].diagonal().add_(2.0)
        rhs = torch.randn((P,), device=DEV, generator=g)
        A0 = A.clone() if P <= 20000 else None
        r = rhs.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = ops.cholesky_solve(A, r)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res = None
        if A0 is not None:
            res = float((A0[:, :P] @ r - rhs).norm() / rhs.norm())
        print(json.dumps({"probe": "chol", "P": P, "ms": ms, "tflops": P ** 3 / 3 / ms / 1e9, "info": int(info.item()), "residual": res}))
if what in ("env",):
    for name, S, rin, f, rout, mk in (("cfg3", 515345, 24, 2, 24, ops.MAP_SINCOS), ("cfg3_ident", 515345, 24, 2, 24, ops.MAP_IDENTITY),
                                      ("cfg4_class", 60000 * 9, 38, 2, 38, ops.MAP_SINCOS), ("cfg5a", 1000000, 38, 29, 38, ops.MAP_IDENTITY),
                                      ("cfg5b", 1000000, 38, 6, 38, ops.MAP_POLY)):
        g = torch.Generator(device=DEV).manual_seed(1)
        env = torch.randn((S, rin), device=DEV, generator=g)
        X = torch.rand((S, 32 if mk != ops.MAP_IDENTITY else f), device=DEV, generator=g)
        core = torch.randn((rin, f, rout), device=DEV, generator=g)
        dot = torch.randn((S, rout), device=DEV, generator=g)
        fx = Factor(X, m=f, map_kind=mk, col=3 if mk != ops.MAP_IDENTITY else 0)
        out = torch.empty((S, rout), device=DEV)
        yh = torch.empty((S,), device=DEV)
        ms_env = timeit(lambda: ops.env_update(env, fx, core, S, out=out), n=5)
        ms_pred = timeit(lambda: ops.predict(env, fx, core, dot, S, out=yh), n=5)
        xb = 8 * (f if mk == ops.MAP_IDENTITY else 1)
        by_env = S * (8.0 * (rin + rout) + xb)
        by_pred = S * (8.0 * (rin + rout) + xb + 8)
        fl = 2.0 * S * rin * f * rout
        print(json.dumps({"probe": "env", "site": name, "rows": S, "ms_env": ms_env, "GBs_env": by_env / ms_env / 1e6, "ms_predict": ms_pred,
                          "GBs_predict": by_pred / ms_pred / 1e6, "tflops_env": fl / ms_env / 1e9}))
