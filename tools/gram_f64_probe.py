"""fp64 Gram (the bit-comparable mode): the factored-operand kernel (gram_f64_fact_kernel, default) against the unfactored
three-barrier kernel (TN_GRAM_F64_UNFACTORED=1) -- correctness on awkward shapes against a torch fp64 einsum and the old kernel,
then timing on the BASELINE sites.  One JSON line per measurement."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
g = torch.Generator(device="cuda").manual_seed(0)
npair = lambda m: m * (m + 1) // 2
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"


def pairs(F):
    m = F.shape[1]
    iu = torch.triu_indices(m, m, device=F.device)
    return F[:, iu[0]] * F[:, iu[1]]


def feat(raw, kind, m):
    if kind == ops.MAP_SINCOS:
        a = (0.5 * 3.14159265358979323846) * raw
        return torch.stack([torch.cos(a), torch.sin(a)], dim=1)
    return torch.stack([raw ** d for d in range(m)], dim=1)


def run(var, args, **kw):
    """fact: the default (16-byte cp.async where the factor rows allow); fact8: 8-byte copies only; old: the unfactored kernel."""
    for k in ("TN_GRAM_F64_UNFACTORED", "TN_GRAM_F64_CP8"):
        os.environ.pop(k, None)
    if var == "old":
        os.environ["TN_GRAM_F64_UNFACTORED"] = "1"
    elif var == "fact8":
        os.environ["TN_GRAM_F64_CP8"] = "1"
    return ops.gram(*args, **kw)


# ---- correctness: ragged row counts, no weights, mapped middle / outer factors, shared rows (div), accumulate, box and linear tilings
cases = [("tiny", 37, 3, 2, 4, {}), ("ragged_box", 1000, 5, 12, 7, {}), ("no_w", 333, 6, 3, 6, {"w": None}),
         ("sincos_mid", 2051, 10, 2, 9, {"map_b": ops.MAP_SINCOS}), ("poly_mid", 777, 8, 4, 8, {"map_b": ops.MAP_POLY}),
         ("poly_box", 4100, 4, 11, 5, {"map_b": ops.MAP_POLY}), ("div_a", 640, 5, 3, 6, {"div_a": 4}), ("one_c", 500, 20, 9, 1, {}),
         ("one_b", 900, 17, 1, 13, {}), ("acc", 70000, 7, 2, 7, {"accumulate": True}), ("wide", 300, 40, 12, 38, {}),
         ("sincos_a", 1500, 2, 5, 6, {"map_a": ops.MAP_SINCOS}),
         # the engine's own role orders (network.py::_roles puts the input factor first): the launcher exchanges the first two factors
         ("swap_lin_cfg3", 3000, 2, 24, 24, {"map_a": ops.MAP_SINCOS}), ("swap_lin_cfg5b", 1200, 6, 38, 38, {}), ("swap_box", 800, 11, 12, 3, {}),
         ("swap_lin_acc", 9000, 3, 20, 5, {"accumulate": True}), ("vec_even", 2100, 38, 6, 38, {}), ("vec_odd_m", 1100, 9, 4, 13, {})]
worst = 0.0
for name, S, ma, mb, mc, o in cases:
    w = None if "w" in o else torch.rand((S,), device="cuda", generator=g) + 0.5
    da = o.get("div_a", 1)
    if "map_a" in o:
        rawa = torch.rand((S,), device="cuda", generator=g) * 2 - 1
        fa, Fa = Factor(rawa.view(S, 1), m=ma, map_kind=o["map_a"]), feat(rawa, o["map_a"], ma)
    else:
        Ta = torch.randn((S // da, ma), device="cuda", generator=g)
        fa, Fa = Factor(Ta, m=ma, div=da), Ta.repeat_interleave(da, dim=0)
    if "map_b" in o:
        X = torch.rand((S, 3), device="cuda", generator=g) * 2 - 1           # the raw column is a strided view, as in MappedInput
        fb, Fb = Factor(X, m=mb, map_kind=o["map_b"], col=1), feat(X[:, 1], o["map_b"], mb)
    else:
        Fb = torch.rand((S, mb), device="cuda", generator=g)
        fb = Factor(Fb, m=mb)
    Fc = torch.randn((S, mc + 3), device="cuda", generator=g)[:, :mc]          # row stride != m
    fc = Factor(Fc, m=mc)
    ref = torch.einsum("s,sa,sb,sc->abc", w if w is not None else torch.ones(S, device="cuda"), pairs(Fa), pairs(Fb), pairs(Fc)).reshape(-1)
    args = (ops.GRAM_FP64, fa, fb, fc, w, S)
    out = {}
    for var in ("fact", "fact8", "old"):
        if o.get("accumulate"):
            M0 = torch.full_like(ref, 0.25)
            out[var] = run(var, args, M=M0.clone(), accumulate=True) - 0.25
        else:
            out[var] = run(var, args)
    torch.cuda.synchronize()
    e_new = max(float((out["fact"] - ref).norm() / ref.norm()), float((out["fact8"] - ref).norm() / ref.norm()))
    e_old = float((out["old"] - ref).norm() / ref.norm())
    worst = max(worst, e_new)
    print(json.dumps({"case": name, "rows": S, "m": [ma, mb, mc], "rel_err_fact_vs_einsum": e_new, "rel_err_old_vs_einsum": e_old,
                      "finite": bool(torch.isfinite(out["fact"]).all())}), flush=True)
print(json.dumps({"worst_rel_err_fact": worst, "ok": worst < 1e-13}), flush=True)

# ---- timing on the BASELINE sites
sites = (("cfg3_mid", 515345, 24, 2, 24, 0), ("cfg3_mid_sincos", 515345, 24, 2, 24, ops.MAP_SINCOS), ("cfg3_engine_order", 515345, 2, 24, 24, 0),
         ("cfg5a_mid", 32768, 38, 29, 38, 0), ("cfg5a_engine_order", 32768, 38, 38, 29, 0),
         ("cfg5b_mid", 131072, 38, 6, 38, 0), ("cfg5b_mid_poly", 131072, 38, 6, 38, ops.MAP_POLY), ("cfg5b_engine_order", 131072, 6, 38, 38, 0),
         ("cfg2_site", 20640, 100, 1, 9, 0), ("cfg1", 4177, 6, 9, 6, 0))
if QUICK:
    sites = (("cfg3_engine_order", 515345, 2, 24, 24, 0), ("cfg5a_engine_order", 8192, 38, 38, 29, 0), ("cfg5b_engine_order", 65536, 6, 38, 38, 0),
             ("cfg2_site", 20640, 100, 1, 9, 0))
for name, S, ma, mb, mc, kind in sites:
    Fa = torch.randn((S, ma), device="cuda", generator=g); Fc = torch.randn((S, mc), device="cuda", generator=g)
    if kind:
        X = torch.rand((S, 4), device="cuda", generator=g) * 2 - 1
        fb = Factor(X, m=mb, map_kind=kind, col=2)
    else:
        fb = Factor(torch.rand((S, mb), device="cuda", generator=g), m=mb)
    w = torch.rand((S,), device="cuda", generator=g) + 0.5
    args = (ops.GRAM_FP64, Factor(Fa, m=ma), fb, Factor(Fc, m=mc), w, S)
    fl = 2.0 * S * npair(ma) * npair(mb) * npair(mc)
    outs = {}
    for var in ("fact", "fact8", "old"):
        M = run(var, args); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run(var, args, M=M)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        outs[var] = M.clone()
        print(json.dumps({"site": name, "rows": S, "kernel": var, "ms": ms, "tflops_fp64": fl / ms / 1e9,
                          "ksplit": int(ops._lib.load().tn_gram_ksplit(S, ma, mb, mc, 0))}), flush=True)
    print(json.dumps({"site": name, "rel_diff": float((outs["fact"] - outs["old"]).norm() / outs["old"].norm())}), flush=True)
