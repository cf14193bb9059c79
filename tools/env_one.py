"""A few launches of the environment kernel at config-3 / config-4 / config-5b shapes (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from tensornetworksfork_b200 import ops
from tensornetworksfork_b200.ops import Factor
g = torch.Generator(device="cuda").manual_seed(0)
for S, r, f, mk in ((515345, 24, 2, ops.MAP_SINCOS), (540000, 38, 2, ops.MAP_SINCOS), (1000000, 38, 6, ops.MAP_POLY)):
    env = torch.randn((S, r), device="cuda", generator=g)
    X = torch.rand((S, 32), device="cuda", generator=g)
    core = torch.randn((r, f, r), device="cuda", generator=g)
    out = torch.empty((S, r), device="cuda")
    for _ in range(2):
        ops.env_update(env, Factor(X, m=f, map_kind=mk, col=5), core, S, out=out)
torch.cuda.synchronize()
print("ok")
