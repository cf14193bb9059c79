"""CPD (canonical polyadic) models on the same sweep engine.

Mirror of the reference's ``CPDNetwork`` (tensor/network.py:935-986): factor j contributes
``Z_j[s,b(,o)] = sum_p x_j[s,p] A_j[b,p(,o)]``; the prediction is ``sum_b prod_j Z_j``; the Jacobian of
factor i is the row-wise Kronecker product of ``prod_{j != i} Z_j`` (S x R) and ``x_i`` (S x f) -- two of
the three factors the Gram kernel takes (the third is the class factor for factor 1, or trivial).
Only the updated factor's Z is refreshed after an update (network.py:976-984).
"""
import torch

from .. import ops
from ..ops import Factor
from .bregman import hessian_terms
from .network import TensorNetwork, MappedInput


class CPDNetwork(TensorNetwork):
    _supports_gradient = False

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._Z = {}

    # -- plan: every factor is (b, p[, o]); factor 0 owns the output leg
    def _plan(self):
        if self._sites is not None:
            return self._sites
        sites = []
        for j, node in enumerate(self.main_nodes):
            labs = list(node.dim_labels)
            if labs not in (["b", "p"], ["b", "p", "o"], ["p", "o"]):
                raise NotImplementedError(f"{node.name}: unsupported CPD factor labels {labs}")
            if ("o" in labs) != (j == 0):
                raise NotImplementedError("the output leg must sit on the first factor")
            sites.append(node)
        self._sites = sites
        return sites

    def _stamp(self):
        return [(id(n.tensor), n.tensor._version) for n in self._plan()]

    def _core_changed(self, k):
        self._Z.pop(k, None)
        self._stamps = self._stamp()

    def _check_external(self):
        st = self._stamp()
        if self._stamps != st:
            self._Z.clear()
            self._stamps = st

    def reset_stacks(self, node=None):
        if node is not None and node in self.main_nodes:
            self._Z.pop(self.main_nodes.index(node), None)
        else:
            self._Z.clear()

    def set_input(self, x):
        key = self._key_of(x)
        if key == self._data_key:
            return False
        self._data_key = key
        self._data = (x,) + self._bind(x)
        self._Z.clear()
        return True

    def _bind(self, x):
        nodes = self._plan()
        facs = []
        if isinstance(x, MappedInput):
            raise NotImplementedError("CPD models take the raw (N, f) matrix")
        for j, node in enumerate(nodes):
            t = x[j] if isinstance(x, (list, tuple)) else x
            if t.dim() != 2 or t.stride(1) != 1:
                t = t.reshape(t.shape[0], -1).contiguous()
            f = node.dim_size("p")
            if t.shape[1] != f:
                raise ValueError(f"factor {j}: input has {t.shape[1]} features, factor expects {f}")
            facs.append(Factor(t, m=f))
        t0 = x[0] if isinstance(x, (list, tuple)) else x
        return facs, t0.shape[0], t0.device

    def _rank(self):
        n0 = self._plan()[0]
        return n0.dim_size("b") if "b" in n0.dim_labels else 1

    def _num_outputs(self):
        return self._plan()[0].dim_size("o")

    def _z(self, j, facs, S, cache):
        """Z_j as (S, R) for j > 0 and (S, R, O) for j = 0."""
        if cache is not None and j in cache:
            return cache[j]
        node = self._plan()[j]
        A = node.tensor
        R = self._rank()
        f = node.dim_size("p")
        if j == 0:
            O = node.dim_size("o")
            A3 = A.reshape(R, f, O).permute(1, 0, 2).reshape(1, f, R * O)
            z = ops.env_update(None, facs[j], A3, S).view(S, R, O)
        else:
            z = ops.env_update(None, facs[j], A.t().reshape(1, f, R), S)
        if cache is not None:
            cache[j] = z
        return z

    def _others(self, i, facs, S, cache):
        """prod_{j != i} Z_j as (S, R); the output leg of Z_0 is summed for i > 0 (network.py:958)."""
        n = len(self._plan())
        prod = None
        for j in range(n):
            if j == i:
                continue
            z = self._z(j, facs, S, cache)
            if j == 0:
                z = z.sum(dim=2)
            prod = z if prod is None else prod * z
        return prod

    def _chain_forward(self, x):
        facs, S, dev = self._bind(x)
        self._require_cuda(dev)
        return self._predict(facs, S, None)

    def _predict(self, facs, S, cache):
        z0 = self._z(0, facs, S, cache)                     # (S, R, O)
        rest = None
        for j in range(1, len(self._plan())):
            z = self._z(j, facs, S, cache)
            rest = z if rest is None else rest * z
        if rest is not None:
            z0 = z0 * rest.unsqueeze(-1)
        return z0.sum(dim=1)

    def _site_problem(self, k, y, loss_fn):
        _, facs, S, dev = self._data
        node = self._plan()[k]
        R = self._rank()
        f = node.dim_size("p")
        O = self._num_outputs()
        yhat = self._predict(facs, S, self._Z)
        if getattr(self, "_yhat_offset", None) is not None:
            yhat = yhat + self._yhat_offset
        loss, g, U, lam = hessian_terms(loss_fn, yhat, y)
        g = g.reshape(S, O).contiguous()
        V = lam.shape[1]
        one = ops.ones_factor(yhat)
        if len(self._plan()) > 1:
            other = self._others(k, facs, S, self._Z).contiguous()
            fo = Factor(other, m=R)
        else:
            other, fo = None, one
        xk = facs[k]
        if k == 0 and O > 1:
            xv = Factor(xk.tensor, m=xk.m, div=V, map_kind=xk.map_kind, col=xk.col)
            fov = Factor(fo.tensor, m=fo.m, div=V if other is not None else fo.div)
            prob = dict(gram=(fov, xv, Factor(U.reshape(S * V, O).contiguous(), m=O)), gw=lam.reshape(S * V).contiguous(),
                        grows=S * V, rhs=(fo, xk, Factor(g, m=O)), rw=None, rrows=S, m_pos=(R, f, O))
        else:
            usum = U.reshape(S, V, O).sum(dim=2)
            w = (lam.reshape(S, V) * usum ** 2).sum(dim=1).contiguous()
            gw = g.sum(dim=1).contiguous()
            prob = dict(gram=(fo, xk, one), gw=w, grows=S, rhs=(fo, xk, one), rw=gw, rrows=S, m_pos=(R, f, 1))
        prob["yhat"] = yhat
        prob["loss"] = loss
        prob["keep"] = (other, U, lam, g)
        return prob

    def _canon(self, k):
        node = self._plan()[k]
        R = self._rank()
        f = node.dim_size("p")
        O = node.dim_size("o") if "o" in node.dim_labels else 1
        return node.tensor.reshape(R, f, O)

    def _layout(self, k):
        shp = list(self._plan()[k].tensor.shape)
        return shp, list(range(len(shp)))

    def _from_canon(self, k, t):
        return t.reshape(self._plan()[k].tensor.shape).contiguous()

    def node_orthonormalize_left(self, node):
        raise NotImplementedError("CPD factors have no bonds to re-gauge")

    node_orthonormalize_right = node_orthonormalize_left
