"""Test infrastructure.  Minimal stand-in for the `opt_einsum` package, which the reference's authors have installed
(environment.yml) and this image lacks: torch.einsum only asks it for `contract_path` (baseline B).

With it importable BEFORE torch, `torch.backends.opt_einsum.is_available()` is True and the three-operand Gram einsum of the
reference (tensor/network.py:212) contracts (J, H) first instead of building the S x P x P temporary of the left-to-right
default.  Nothing of the reference is changed: this only restores its authors' environment.
"""
__version__ = "3.3.0"


def contract_path(*args, **kwargs):
    n = sum(1 for a in args[1:] if hasattr(a, "shape")) if isinstance(args[0], str) else len(args) // 2
    # three operands = the Gram einsum (J*, J, H): (J H) first, as opt_einsum chooses; otherwise left to right, pairwise
    path = [(1, 2), (0, 1)] if n == 3 else [(0, 1)] * max(n - 1, 1)
    return path, None
