"""Cum-sum tensor train: every monomial x_{p1} x_{p2} ... x_{pn} with p1 <= p2 <= ... <= pn counted once.

The reference builds this model by inserting a dense (f, f, f, f) operator node ``C = H.D`` between each input and
each core (``CumSumLayer`` / ``get_cum_sum_operator``, tensor/layers.py:408-477) and contracting it with einsum.  Its
closed form (verified against the reference) is a running sum over the feature index:

    E_1[s,p,b] = x[s,p] G_1[p,b]                                     Lc_k = prefix-sum_p E_k
    E_k[s,p,b] = x[s,p] sum_a Lc_{k-1}[s,p,a] G_k[a,p,b]
    T_k[s,p,a] = x[s,p] sum_b G_k[a,p,b] Rc_{k+1}[s,p,b]             Rc_k = suffix-sum_p T_k
    J_k[s,(a,p,b)] = x[s,p] Lc_{k-1}[s,p,a] Rc_{k+1}[s,p,b]          yhat = sum_i J_k[s,i] theta_k[i]

Environments are (S, f, r), one small per-feature contraction per p (the ordinary environment kernel on strided views);
the Jacobian couples a, b through p, so it is not a Kronecker product and its Gram goes through the table-driven
kernel ``tn_gram_generic``.  Single output (C = 1) only.
"""
import torch

from .. import ops
from ..ops import Factor
from .bregman import hessian_terms
from .network import TensorNetwork, MappedInput, batch_mean_of_means


class CumSumNetwork(TensorNetwork):
    _supports_gradient = False

    def _plan(self):
        sites = super()._plan()
        for s in sites:
            if s.cls is not None and s.node.dim_size(s.cls) != 1:
                raise NotImplementedError("cum-sum trains with more than one output")
            if isinstance(s.phys, tuple):
                raise NotImplementedError("cum-sum trains have one input per core")
        return sites

    def _phys_behind_operator(self, node):
        """The reference's CumSumLayer graph (layers.py:435-457): core -[p_k]- operator node O_k -[d_k]- input.  The dense operator
        is not contracted here (closed form), so the core's physical leg is bound to the input behind the operator."""
        for lab in node.dim_labels:
            op = node.connections.get(lab)
            if op is None or node.is_horizontal_bond(lab):
                continue
            far = [n2 for l2, n2 in op.connections.items() if l2 != lab and any(n2 is n for n in self.input_nodes)]
            if len(far) == 1:
                return [(lab, far[0])]
        return []

    def _bind(self, x):
        if isinstance(x, (MappedInput, list, tuple)):
            raise NotImplementedError("cum-sum trains take one (N, f) matrix shared by all cores (poly-mode)")
        return super()._bind(x)

    # -- per-feature environment steps on (S, f, r) tensors
    def _xcol(self, fac, p):
        return Factor(fac.tensor, m=1, col=p)

    def _left_step(self, prev, fac, k, S):
        G = self._canon(k)                       # (rl, 1, f, rr)
        rl, _, f, rr = G.shape
        E = torch.empty((S, f, rr), dtype=torch.float64, device=G.device)
        for p in range(f):
            ops.env_update(None if prev is None else prev[:, p, :], self._xcol(fac, p), G[:, 0, p, :].reshape(rl, 1, rr), S,
                           out=E[:, p, :])
        return torch.cumsum(E, dim=1)

    def _right_step(self, nxt, fac, k, S):
        G = self._canon(k)
        rl, _, f, rr = G.shape
        T = torch.empty((S, f, rl), dtype=torch.float64, device=G.device)
        for p in range(f):
            ops.env_update(None if nxt is None else nxt[:, p, :], self._xcol(fac, p), G[:, 0, p, :].t().reshape(rr, 1, rl), S,
                           out=T[:, p, :])
        return torch.flip(torch.cumsum(torch.flip(T, dims=[1]), dim=1), dims=[1])

    def _get_left(self, k):
        if k < 0:
            return None
        _, facs, S, _ = self._data
        j = k
        while j >= 0 and j not in self._left:
            j -= 1
        env = self._left[j] if j >= 0 else None
        for i in range(j + 1, k + 1):
            env = self._left_step(env, facs[i], i, S)
            self._left[i] = env
        return env

    def _get_right(self, k):
        n = len(self._plan())
        if k >= n:
            return None
        _, facs, S, _ = self._data
        j = k
        while j < n and j not in self._right:
            j += 1
        env = self._right[j] if j < n else None
        for i in range(j - 1, k - 1, -1):
            env = self._right_step(env, facs[i], i, S)
            self._right[i] = env
        return env

    def _predict_at(self, k, L, R, fac, S, G=None):
        """yhat[s] = sum_p x_p sum_{a,b} Lc[s,p,a] G_k[a,p,b] Rc[s,p,b]; ``G`` (rl, 1, f, rr) stands in for core k (the J v pass
        of the matrix-free sweeps: the prediction is linear in the core)."""
        G = self._canon(k) if G is None else G
        rl, _, f, rr = G.shape
        dev = G.device
        one = torch.ones((1, 1), dtype=torch.float64, device=dev)
        acc = torch.zeros((S,), dtype=torch.float64, device=dev)
        tmp = torch.empty((S,), dtype=torch.float64, device=dev)
        for p in range(f):
            dot = one if R is None else R[:, p, :]
            ops.predict(None if L is None else L[:, p, :], self._xcol(fac, p), G[:, 0, p, :].reshape(rl, 1, rr), dot, S,
                        dot_div=(1 << 30) if R is None else 1, out=tmp)
            acc += tmp
        return acc.view(S, 1)

    def _chain_forward(self, x):
        facs, S, dev = self._bind(x)
        self._require_cuda(dev)
        n = len(self._plan())
        env = None
        for k in range(n - 1):
            env = self._left_step(env, facs[k], k, S)
        return self._predict_at(n - 1, env, None, facs[n - 1], S)

    def _tables(self, rl, f, rr, has_l, has_r, dev):
        a = torch.arange(rl, device=dev).view(rl, 1, 1).expand(rl, f, rr)
        p = torch.arange(f, device=dev).view(1, f, 1).expand(rl, f, rr)
        b = torch.arange(rr, device=dev).view(1, 1, rr).expand(rl, f, rr)
        t1 = (p * rl + a) if has_l else torch.zeros_like(p)
        t3 = (p * rr + b) if has_r else torch.zeros_like(p)
        return [t.reshape(-1).to(torch.int32).contiguous() for t in (t1, p, t3)]

    def _one_update(self, k, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss):
        self._check_external()
        _, facs, S, dev = self._data
        G = self._canon(k)
        rl, _, f, rr = G.shape
        L = self._get_left(k - 1)
        R = self._get_right(k + 1)
        yhat = self._predict_at(k, L, R, facs[k], S)
        if getattr(self, "_yhat_offset", None) is not None:
            yhat = yhat + self._yhat_offset
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        loss, g, U, lam = hessian_terms(loss_fn, yhat if out_labels else yhat[:, 0], y)
        V = lam.shape[1]
        w = (lam.reshape(S, V) * U.reshape(S, V) ** 2).sum(dim=1).contiguous()
        gw = g.reshape(S).contiguous()
        one = ops.ones_factor(G)
        f1 = one if L is None else Factor(L.reshape(S, f * rl), m=f * rl)
        f3 = one if R is None else Factor(R.reshape(S, f * rr), m=f * rr)
        t1, t2, t3 = self._tables(rl, f, rr, L is not None, R is not None, dev)
        P = rl * f * rr
        buf = torch.empty((P * P + P,), dtype=torch.float64, device=dev)
        A, b = buf[:P * P], buf[P * P:]
        ops.gram_generic(f1, facs[k], f3, t1, t2, t3, w, S, out=A)
        ops.gram_generic(f1, facs[k], f3, t1, t2, t3, gw, S, rhs_only=True, out=b)
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(buf, group=self.process_group)
        node = self.main_nodes[k]
        theta = _Theta(G.contiguous().view(-1))
        step_c = self.solve_system(theta, A.view(P, P), b, method=method, eps=eps)
        step = self._from_canon(k, step_c.reshape(G.shape))
        new = node.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.contiguous().view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
        node.tensor = new
        self._core_changed(k)
        if not need_loss:
            return None
        return batch_mean_of_means(loss, batch_size, row_offset=self.shard_offset,
                                   n_total=self.shard_total if self.process_group is not None else S, group=self.process_group)

    def node_orthonormalize_left(self, node):
        raise NotImplementedError("QR re-gauge is not defined for the cum-sum train (the operator couples the bonds)")

    node_orthonormalize_right = node_orthonormalize_left

    def _krylov_problem(self, node, y, loss_fn):
        """(per-row loss, b, matvec) of one core for ``lanczos_swipe`` / ``scipy_swipe`` (the reference runs them on the operator-node
        graph, network.py:709-932).  J v is the prediction with v in place of the core, J^T u the table-driven right-hand-side
        pass with row weights u; everything flat in canonical (a, p, b) order."""
        k = self.main_nodes.index(node)
        _, facs, S, dev = self._data
        G = self._canon(k)
        rl, _, f, rr = G.shape
        L = self._get_left(k - 1)
        R = self._get_right(k + 1)
        yhat = self._predict_at(k, L, R, facs[k], S)
        if getattr(self, "_yhat_offset", None) is not None:
            yhat = yhat + self._yhat_offset
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        loss, g, U, lam = hessian_terms(loss_fn, yhat if out_labels else yhat[:, 0], y)
        V = lam.shape[1]
        w = (lam.reshape(S, V) * U.reshape(S, V) ** 2).sum(dim=1).contiguous()
        one = ops.ones_factor(G)
        f1 = one if L is None else Factor(L.reshape(S, f * rl), m=f * rl)
        f3 = one if R is None else Factor(R.reshape(S, f * rr), m=f * rr)
        t1, t2, t3 = self._tables(rl, f, rr, L is not None, R is not None, dev)
        P = rl * f * rr
        group = self.process_group

        def jt(u):
            out = torch.empty((P,), dtype=torch.float64, device=dev)
            ops.gram_generic(f1, facs[k], f3, t1, t2, t3, u.contiguous(), S, rhs_only=True, out=out)
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(out, group=group)
            return out

        def matvec(v):
            jv = self._predict_at(k, L, R, facs[k], S, G=v.contiguous().view(rl, 1, f, rr))
            return jt(w * jv.view(S))

        return loss, jt(g.reshape(S)), matvec


class _Theta:
    """Minimal stand-in for a node: solve_system only reads ``.tensor`` (canonical order here)."""

    def __init__(self, t):
        self.tensor = t
