"""Margins of the BASELINE-size recordings in both Gram modes (GPU): per-update relative loss errors, prediction and core errors."""
import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
torch.set_default_dtype(torch.float64)
import cfg1_case, cfg2_case, cfg5b_case
np.set_printoptions(precision=2, linewidth=250)
for gm in ("fp64", "tf32x3", "tf32", "f16"):
    le, pe, ce = cfg1_case.run("cuda", gram_mode=gm); print("cfg1", gm, "loss_err", le, "pred", pe, "core", ce, flush=True)
    le, pe = cfg2_case.run("cuda", gram_mode=gm); print("cfg2", gm, "loss_err", le, "pred", pe, flush=True)
    le, pe = cfg5b_case.run("cuda", gram_mode=gm); print("cfg5b", gm, "loss_err", le, "pred", pe, flush=True)
