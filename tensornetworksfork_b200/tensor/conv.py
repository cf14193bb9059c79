"""Patch/pixel "conv-TT" network: the engine behind ``TensorConvolutionTrainLayer``.

Reference: tensor/layers.py:791-890 builds, per column k, an input node ``x[s, patches, patch_pixels]`` (the same tensor for
every column), a patch core ``A_k[r_k, (c), patches, r_{k+1}]`` and a pixel core ``C_k[CB_k, patch_pixels, CB_{k+1}]``; the
train-node order is ``C1, A1, C2, A2, ...`` and the generic sweep code of tensor/network.py contracts the graph by labels.
The scripts that use it (image_convolution_CG_MNIST.py:95, image_convolution_lanczos_MNIST.py:92, CIFAR_minres.py) drive it
with the matrix-free sweeps ``scipy_swipe`` / ``lanczos_swipe``, because the patch cores are large (P = r*patches*r).

Closed forms used here (q = patch, t = pixel, alpha/beta = pixel-core bonds, c = output leg):

    Y_k[s,q,a,b]   = sum_t x[s,q,t] C_k[a,t,b]
    E_k[s,c,b,r']  = sum_{r,a,q} E_{k-1}[s,c,a,r] A_k[r,q,r'] Y_k[s,q,a,b]            (left environment, E_1 from A_1 alone)
    R_k[s,a,r]     = sum_{r',b,q} A_k[r,q,r'] Y_k[s,q,a,b] R_{k+1}[s,b,r']            (right environment)
    yhat[s,c]      = E_n[s,c,0,0]
    J_{A_k}[s,c,(r,q,r')] = sum_a E_{k-1}[s,c,a,r] * YR_a[s,(q,r')],   YR_a[s,q,r'] = sum_b Y_k[s,q,a,b] R_{k+1}[s,b,r']
    J_{C_k}[s,c,(a,t,b)]  = sum_q K[s,c,a,b,q] x[s,q,t],               K = sum_{r,r'} E_{k-1}[s,c,a,r] A_k[r,q,r'] R_{k+1}[s,b,r']

The Jacobian of a patch core is a sum of CB Kronecker terms and is never formed: its matvec is CB prediction passes and CB
right-hand-side passes of the same kernels the tensor-train path uses.  The Jacobian of a pixel core is small (P <= CB*pixels*CB)
and is materialised per row chunk.  All arithmetic runs in libtn_b200.so through ``ops``; there is no CPU path.
"""
import torch

from .. import ops
from ..ops import Factor
from .bregman import hessian_terms
from .network import TensorNetwork


class ConvTrainNetwork(TensorNetwork):
    _supports_gradient = False

    def __init__(self, input_nodes, main_nodes, train_nodes=None, output_labels=("s",), sample_dim="s"):
        super().__init__(input_nodes, main_nodes, train_nodes, output_labels=output_labels, sample_dim=sample_dim)
        self.chunk_rows = 16384         # rows processed at a time (bounds the per-chunk temporaries)
        self._cols = None
        self._cache = {}
        self._ver = {}                  # node -> number of updates applied by this engine (part of every cache stamp)
        self.matvec_count = 0           # matvecs served so far (bench bookkeeping)
        self.dense_chunk_bytes = 1 << 30   # cap on the materialised Jacobian of one chunk in the dense sweep

    # ------------------------------------------------------------------ graph
    def _columns(self):
        """[(A_k, C_k)] per column, recognised from the node graph."""
        if self._cols is not None:
            return self._cols
        cols = []
        for A in self.main_nodes:
            if "patches" not in A.connections:
                raise NotImplementedError(f"{A.name}: not a patch core of a TensorConvolutionTrainLayer")
            xn = A.connections["patches"]
            if "patch_pixels" not in xn.connections:
                raise NotImplementedError(f"{xn.name}: input node without a pixel core")
            Cn = xn.connections["patch_pixels"]        # a plain pixel vector (one column, or convolution_bond <= 0) is a core with bonds 1
            cols.append((A, Cn))
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        owners = [k for k, (A, _) in enumerate(cols) if any(l in out_labels for l in A.dim_labels)]
        if owners not in ([], [0]):
            raise NotImplementedError("the output leg must sit on the first patch core")
        self._cols = cols
        return cols

    def _plan(self):
        raise NotImplementedError("ConvTrainNetwork keeps its own column plan")

    @staticmethod
    def _to_canon(node, order):
        present = [l for l in order if l in node.dim_labels]
        if present != list(node.dim_labels):
            raise NotImplementedError(f"{node.name}: label order {node.dim_labels} is not the constructor's")
        # grow_cart leaves the old last cores as stride-0 broadcasts until their first update: the kernels read dense rows
        t = node.tensor if node.tensor.is_contiguous() else node.tensor.contiguous()
        return t.reshape([node.dim_size(l) if l in node.dim_labels else 1 for l in order])

    def _bonds(self, k, which):
        """(left, right) bond labels of the patch (which = 0) or pixel (which = 1) core of column k, read from the connections to the
        neighbouring columns: grow_cart gives the old last cores a right bond without listing it in ``right_labels``."""
        cols = self._columns()
        node = cols[k][which]

        def towards(j):
            if not 0 <= j < len(cols):
                return None
            return next((lab for lab, other in node.connections.items() if other is cols[j][which] and lab in node.dim_labels), None)
        return towards(k - 1) or "_l", towards(k + 1) or "_r"

    def _A4(self, k):
        """Patch core k as (r, c, Q, r')."""
        A = self._columns()[k][0]
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        cls = next((l for l in A.dim_labels if l in out_labels), "_c")
        left, right = self._bonds(k, 0)
        return self._to_canon(A, [left, cls, "patches", right])

    def _C3(self, k):
        """Pixel core k as (a, T, a')."""
        Cn = self._columns()[k][1]
        left, right = self._bonds(k, 1)
        return self._to_canon(Cn, [left, "patch_pixels", right])

    def _num_outputs(self):
        return self._A4(0).shape[1]

    def _locate(self, node):
        for k, (A, Cn) in enumerate(self._columns()):
            if node is A:
                return "A", k
            if node is Cn:
                return "C", k
        raise ValueError(f"{node.name} is not a node of this network")

    # ------------------------------------------------------------------ data / caches
    def _bind(self, x):
        if isinstance(x, (list, tuple)):
            if not all(t is x[0] for t in x):
                raise NotImplementedError("conv-TT columns share one input tensor")
            x = x[0]
        if x.dim() != 3:
            raise ValueError(f"expected x of shape (N, patches, patch_pixels), got {tuple(x.shape)}")
        A4, C3 = self._A4(0), self._C3(0)
        if x.shape[1] != A4.shape[2] or x.shape[2] != C3.shape[1]:
            raise ValueError(f"x has {x.shape[1]} patches x {x.shape[2]} pixels, the cores expect {A4.shape[2]} x {C3.shape[1]}")
        if x.dtype != torch.float64:
            x = x.to(torch.float64)
        return x.contiguous()

    def set_input(self, x):
        key = self._key_of(x)
        if key == self._data_key:
            return False
        self._data_key = key
        xb = self._bind(x)
        self._data = (x, xb, xb.shape[0], xb.device)
        self._cache = {}
        return True

    def reset_stacks(self, node=None):
        self._cache = {}
        self.left_stacks = None
        self.right_stacks = None

    def _stamp_nodes(self, nodes):
        # own update counter (ids of freed tensors can be reused) + identity/version to notice changes made from outside
        return tuple((self._ver.get(n, 0), id(n.tensor), n.tensor._version) for n in nodes)

    def _check_external(self):
        pass    # every cached quantity carries the stamps of the cores it was built from

    def _stamp(self):
        return self._stamp_nodes(self.train_nodes)

    def _with_offset(self, yhat, lo, hi):
        """Add the outputs of the other members of a SumOfNetworks (held fixed while a node of this member is updated)."""
        off = getattr(self, "_yhat_offset", None)
        return yhat if off is None else yhat + off[lo:hi]

    def _require_cuda(self, dev):
        if dev.type != "cuda":
            raise RuntimeError("tensornetworksfork_b200 runs on CUDA devices only (no CPU fallback); move data and "
                               "model with .to('cuda')")
        for n in self.train_nodes:
            if n.tensor.device != dev:
                raise RuntimeError(f"core {n.name} lives on {n.tensor.device}, data on {dev}")

    def _cached(self, key, deps, build):
        st = self._stamp_nodes(deps)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == st:
            return hit[1]
        val = build()
        self._cache[key] = (st, val)
        return val

    def _chunks(self, S):
        step = max(int(self.chunk_rows), 1)
        return [(lo, min(lo + step, S)) for lo in range(0, S, step)]

    # ------------------------------------------------------------------ per-chunk pieces (xc: (s, Q, T) contiguous)
    def _Yt(self, k, xc, tag):
        """Y_k laid out (a, a', s, Q): one contiguous (s, Q) site-input matrix per pixel-bond pair."""
        Cn = self._columns()[k][1]

        def build():
            C3 = self._C3(k)
            a, T, a2 = C3.shape
            s, Q, _ = xc.shape
            Y = ops.rows_dot(xc.view(s * Q, T), C3.permute(0, 2, 1).reshape(a * a2, T).contiguous())      # (s*Q, a*a')
            return Y.view(s, Q, a, a2).permute(2, 3, 0, 1).contiguous()
        return self._cached(("Y", k, tag), [Cn], build)

    @staticmethod
    def _AY(Yt_ab, Aperm, r, r2):
        """AY[s, r, r'] = sum_q Y[s, q] A[r, q, r'] for one pixel-bond pair: the column's per-sample transfer block, contracted over
        the patches FIRST (shared by all output legs) -- one tall-skinny product with the core as the shared matrix."""
        return ops.rows_dot(Yt_ab, Aperm).view(-1, r, r2)

    def _env_left(self, k, xc, tag):
        """E_k (s, C, a', r') after columns 0..k."""
        cols = self._columns()
        deps = [n for j in range(k + 1) for n in cols[j]]

        def build():
            Yt = self._Yt(k, xc, tag)
            A4 = self._A4(k)
            r, c, Q, r2 = A4.shape
            a, a2, s, _ = Yt.shape
            if k == 0:
                E = torch.empty((s, c, a2, r2), dtype=torch.float64, device=xc.device)
                V = A4[0].permute(0, 2, 1).reshape(c * r2, Q).contiguous()           # ((c, r'), q)
                for b in range(a2):
                    E[:, :, b, :] = ops.rows_dot(Yt[0, b], V).view(s, c, r2)
                return E
            Ep = self._env_left(k - 1, xc, tag)                       # (s, C, a, r)
            C = Ep.shape[1]
            E = torch.empty((s, C, a2, r2), dtype=torch.float64, device=xc.device)
            Aperm = A4[:, 0].permute(0, 2, 1).reshape(r * r2, Q).contiguous()         # ((r, r'), q)
            acc = torch.empty((s, C, r2), dtype=torch.float64, device=xc.device)
            for b in range(a2):
                for al in range(a):
                    ops.bmm(Ep[:, :, al, :], self._AY(Yt[al, b], Aperm, r, r2), out=acc, accumulate=al > 0)
                E[:, :, b, :] = acc
            return E
        return self._cached(("L", k, tag), deps, build)

    def _env_right(self, k, xc, tag):
        """R_k (s, a, r) of columns k..n-1."""
        cols = self._columns()
        n = len(cols)
        deps = [nd for j in range(k, n) for nd in cols[j]]

        def build():
            Yt = self._Yt(k, xc, tag)
            A4 = self._A4(k)
            r, _, Q, r2 = A4.shape
            a, a2, s, _ = Yt.shape
            R = torch.empty((s, a, r), dtype=torch.float64, device=xc.device)
            if k == n - 1:
                V = A4[:, 0, :, 0].contiguous()                                      # (r, q)
                for al in range(a):
                    R[:, al, :] = ops.rows_dot(Yt[al, 0], V)
                return R
            Rn = self._env_right(k + 1, xc, tag)                      # (s, a', r')
            Aperm = A4[:, 0].permute(0, 2, 1).reshape(r * r2, Q).contiguous()
            acc = torch.empty((s, r, 1), dtype=torch.float64, device=xc.device)
            for al in range(a):
                for b in range(a2):
                    ops.bmm(self._AY(Yt[al, b], Aperm, r, r2), Rn[:, b, :].unsqueeze(-1), out=acc, accumulate=b > 0)
                R[:, al, :] = acc.view(s, r)
            return R
        return self._cached(("R", k, tag), deps, build)

    def _predict_chunk(self, xc, tag):
        n = len(self._columns())
        return self._env_left(n - 1, xc, tag)[:, :, 0, 0]

    # ------------------------------------------------------------------ forward
    def _chain_forward(self, x):
        xb = self._bind(x)
        self._require_cuda(xb.device)
        saved = self._cache
        outs = []
        try:
            for lo, hi in self._chunks(xb.shape[0]):
                self._cache = {}
                outs.append(self._predict_chunk(xb[lo:hi], "fwd").clone())
        finally:
            self._cache = saved
        return torch.cat(outs, dim=0) if len(outs) != 1 else outs[0]

    def forward_batch(self, x, batch_size):
        return self.forward(x, to_tensor=True)       # batching only bounds memory; _chain_forward already chunks the rows

    # ------------------------------------------------------------------ local problems
    def _loss_terms(self, yhat, y, loss_fn):
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        y_in = yhat if out_labels else yhat[:, 0]
        loss, g, U, lam = hessian_terms(loss_fn, y_in, y)
        S, C = yhat.shape
        return loss, g.reshape(S, C).contiguous(), U.contiguous(), lam.contiguous()

    @staticmethod
    def _apply_H(U, lam, t):
        """u[s,:] = H_s t[s,:] with H_s = sum_v lam[s,v] U[s,v,:] U[s,v,:]^T (C x C per sample; C <= 10)."""
        coef = torch.einsum("svc,sc->sv", U, t) * lam
        return torch.einsum("sv,svc->sc", coef, U).contiguous()

    def _node_chunk(self, kind, k, xc, tag):
        """Operators of node (kind, k) on one row chunk: returns (rhs_fn(g) -> b_flat, jv_fn(v) -> (s,C), jt_fn(u) -> flat)."""
        cols = self._columns()
        n = len(cols)
        dev = xc.device
        s = xc.shape[0]
        one = ops.ones_factor(xc)
        ones11 = torch.ones((1, 1), dtype=torch.float64, device=dev)
        A4 = self._A4(k)
        r, c, Q, r2 = A4.shape
        C = self._num_outputs()
        Yt = self._Yt(k, xc, tag)
        a, a2 = Yt.shape[0], Yt.shape[1]
        Rn = self._env_right(k + 1, xc, tag) if k < n - 1 else None                   # (s, a', r')
        Ep = self._env_left(k - 1, xc, tag) if k > 0 else None                        # (s, C, a, r)

        def wide_rhs(G, W, w=None):
            """(ra, m) flat = sum_rows w[row] G[row,:]^T W[row,:]: the right-hand-side / J^T reduction over the rows."""
            return ops.outer_rows(G, W, w).view(-1)

        def wide_dot(W, vmat):
            """z (rows, ra) = W (rows, m) @ vmat (ra, m)^T: the J v pass over a wide per-row operand."""
            return ops.rows_dot(W, vmat.contiguous())

        if kind == "A":
            # YR_a (s, Q*r'): the site input of the patch core with the right environment folded in
            YR = []
            for al in range(a):
                if Rn is None:
                    YR.append(Yt[al, 0])                                          # (s, Q), r' = 1
                else:
                    Ysl = Yt[al].permute(1, 2, 0)                                 # (s, Q, a') view of (a', s, Q)
                    YR.append(ops.bmm(Ysl, Rn).view(s, Q * r2))
            m = Q * r2
            if k == 0:
                def rhs_fn(g):
                    return wide_rhs(g.contiguous(), YR[0])

                def jv_fn(v):
                    return wide_dot(YR[0], v.view(C, m))
                return rhs_fn, jv_fn, rhs_fn
            Ea = [Ep[:, :, al, :].contiguous() for al in range(a)]               # (s, C, r) each
            eye3 = torch.eye(r, dtype=torch.float64, device=dev).view(r, 1, r)

            def fold(w):
                out = None
                for al in range(a):
                    _, G = ops.class_rows(Ea[al], None, w)                        # (s, r) = sum_c w[s,c] E_a[s,c,:]
                    o = wide_rhs(G, YR[al])
                    out = o if out is None else out.add_(o)
                return out

            def jv_fn(v):
                vm = v.view(r, m)
                t = None
                for al in range(a):
                    z = wide_dot(YR[al], vm)                                      # (s, r)
                    o = ops.predict(Ea[al].view(s * C, r), one, eye3, z, s * C, cdiv=1 << 30, dot_div=C)
                    t = o if t is None else t.add_(o)
                return t.view(s, C)
            return fold, jv_fn, fold

        # ---- pixel core: its Jacobian is small, materialise it
        J = self._jacobian_chunk(kind, k, xc, tag)
        P = J.shape[1]

        def rhs_fn(g):
            return wide_rhs(g.reshape(s * C, 1).contiguous(), J)

        def jv_fn(v):
            return wide_dot(J, v.view(1, P)).view(s, C)
        return rhs_fn, jv_fn, rhs_fn

    def _jacobian_chunk(self, kind, k, xc, tag):
        """Dense Jacobian J (s*C, P) of node (kind, k) on one row chunk, P in the node's own (canonical) order."""
        cols = self._columns()
        n = len(cols)
        dev = xc.device
        s = xc.shape[0]
        A4 = self._A4(k)
        r, c, Q, r2 = A4.shape
        C = self._num_outputs()
        Yt = self._Yt(k, xc, tag)
        a, a2 = Yt.shape[0], Yt.shape[1]
        Rn = self._env_right(k + 1, xc, tag) if k < n - 1 else None
        Ep = self._env_left(k - 1, xc, tag) if k > 0 else None
        if kind == "A":
            m = Q * r2
            YR = torch.empty((s, a, m), dtype=torch.float64, device=dev)
            for al in range(a):
                if Rn is None:
                    YR[:, al, :] = Yt[al, 0]
                else:
                    YR[:, al, :] = ops.bmm(Yt[al].permute(1, 2, 0), Rn).view(s, m)
            if k == 0:
                # the core owns the output leg: J[s, c, (c', q, r')] = delta_cc' YR[s, (q, r')]
                J = torch.zeros((s, C, C, m), dtype=torch.float64, device=dev)
                for cc in range(C):
                    J[:, cc, cc, :] = YR[:, 0, :]
                return J.view(s * C, C * m)
            # J[s, (c, r), (q, r')] = sum_a E[s, c, a, r] YR[s, a, (q, r')]: one small product per sample
            Et = Ep.permute(0, 1, 3, 2).contiguous().view(s, C * r, a)
            return ops.bmm(Et, YR).view(s * C, r * m)
        T = xc.shape[2]
        core_A = A4[:, 0]                                                          # (r, Q, r')
        if k == 0 and Rn is None:
            # a single column: J[s, c, t] = sum_q A[c, q] x[s, q, t], one tall-skinny product with the core as the shared matrix
            Jt = ops.rows_dot(xc.transpose(1, 2).contiguous().view(s * T, Q), A4[0, :, :, 0].contiguous())     # (s*T, c)
            return Jt.view(s, T, c).permute(0, 2, 1).contiguous().view(s * C, T)
        if k == 0:
            K = ops.rows_dot(Rn.view(s * a2, r2), A4[0].reshape(c * Q, r2))
            K = K.view(s, a2, c, Q).permute(0, 2, 1, 3).contiguous().view(s, c * a2, Q)          # (s, (c,b), q), a = 1
        elif Rn is None:
            K = ops.rows_dot(Ep.view(s * C * a, r), core_A[:, :, 0].t().contiguous()).view(s, C * a, Q)       # a' = 1
        else:
            # contract the right environment with the core first (no output leg there), then the left one per sample
            AR = ops.rows_dot(Rn.view(s * a2, r2), core_A.reshape(r * Q, r2)).view(s, a2, r, Q)
            Ks = [ops.bmm(Ep.view(s, C * a, r), AR[:, b]) for b in range(a2)]
            K = torch.stack(Ks, dim=2).view(s, C * a * a2, Q)                      # (s, (c,a,b), q)
        J = ops.bmm(K, xc)                                                         # (s, (c,a,b), T)
        P = a * T * a2
        return J.view(s, C, a, a2, T).permute(0, 1, 2, 4, 3).contiguous().view(s * C, P)

    def _krylov_problem(self, node, y, loss_fn):
        kind, k = self._locate(node)
        _, xb, S, dev = self._data
        parts = []
        b = None
        losses = []
        for ci, (lo, hi) in enumerate(self._chunks(S)):
            xc = xb[lo:hi]
            yhat = self._with_offset(self._predict_chunk(xc, ci), lo, hi)
            loss, g, U, lam = self._loss_terms(yhat, y[lo:hi], loss_fn)
            rhs_fn, jv_fn, jt_fn = self._node_chunk(kind, k, xc, ci)
            bc = rhs_fn(g)
            b = bc if b is None else b.add_(bc)
            parts.append((jv_fn, jt_fn, U, lam))
            losses.append(loss)
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(b, group=self.process_group)

        def matvec(v):
            self.matvec_count += 1
            v = v.contiguous().view(-1)
            out = None
            for jv_fn, jt_fn, U, lam in parts:
                o = jt_fn(self._apply_H(U, lam, jv_fn(v)))
                out = o if out is None else out.add_(o)
            if self.process_group is not None:
                import torch.distributed as dist
                dist.all_reduce(out, group=self.process_group)
            return out

        return torch.cat(losses, dim=0) if len(losses) != 1 else losses[0], b, matvec

    def _apply_step(self, node, step_c, lr):
        new = node.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step_c.contiguous().view(-1), lr=lr)
        node.tensor = new
        self._ver[node] = self._ver.get(node, 0) + 1     # every cached quantity that depends on this core is rebuilt on use

    # ------------------------------------------------------------------ dense sweep (accumulating_swipe)
    def _dense_A_b(self, node, y, loss_fn):
        """A = sum J^T H J (P x P, dense), b = sum J^T g and the per-row loss, accumulated chunk by chunk (reference
        tensor/network.py:174-217).  The Jacobian of a conv-TT core is a sum of Kronecker terms, not one, so it is materialised per
        chunk and reduced with the row-reduced outer-product kernel; the matrix-free sweeps are the fast path for large patch cores."""
        kind, k = self._locate(node)
        _, xb, S, dev = self._data
        P = node.tensor.numel()
        C = self._num_outputs()
        rows_cap = max(1, int(self.dense_chunk_bytes // (8 * C * P)))
        A = torch.zeros((P, P), dtype=torch.float64, device=dev)
        b = torch.zeros((1, P), dtype=torch.float64, device=dev)
        losses = []
        saved_chunk = self.chunk_rows
        self.chunk_rows = min(self.chunk_rows, rows_cap)
        try:
            for ci, (lo, hi) in enumerate(self._chunks(S)):
                xc = xb[lo:hi]
                s = hi - lo
                tag = ("dense", ci)
                yhat = self._with_offset(self._predict_chunk(xc, tag), lo, hi)
                loss, g, U, lam = self._loss_terms(yhat, y[lo:hi], loss_fn)
                J = self._jacobian_chunk(kind, k, xc, tag).view(s, C, P)
                F, Gr = ops.class_rows(J, U, g)                      # virtual rows F (s*V, P), G (s, P) = sum_c g J
                w = lam.reshape(-1).contiguous()
                for i0 in range(0, P, 128):
                    i1 = min(i0 + 128, P)
                    ops.outer_rows(F[:, i0:i1], F, w, out=A[i0:i1], accumulate=True)
                ops.outer_rows(torch.ones((s, 1), dtype=torch.float64, device=dev), Gr, None, out=b, accumulate=True)
                losses.append(loss)
                for key in [kk for kk in self._cache if kk[-1] == tag]:      # dense chunks are not revisited: free their caches
                    del self._cache[key]
        finally:
            self.chunk_rows = saved_chunk
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(A, group=self.process_group)
            dist.all_reduce(b, group=self.process_group)
        return A, b.view(-1), (torch.cat(losses, dim=0) if len(losses) != 1 else losses[0])

    def get_A_b(self, node, grad=None, hessian=None, method=None, y=None, loss_fn=None):
        if loss_fn is None:
            from .network import _FixedTerms
            loss_fn = _FixedTerms(grad, hessian)
        A, b, _ = self._dense_A_b(node, y, loss_fn)
        shp = tuple(node.tensor.shape)
        return A.reshape(shp + shp), b.reshape(shp)

    def _update_node(self, node, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss):
        A, b, loss_rows = self._dense_A_b(node, y, loss_fn)
        step = self.solve_system(node, A, b, method=method, eps=eps)
        new = node.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.contiguous().view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
        node.tensor = new
        self._ver[node] = self._ver.get(node, 0) + 1
        if not need_loss:
            return None
        from .network import batch_mean_of_means
        S = loss_rows.shape[0]
        return batch_mean_of_means(loss_rows, batch_size, row_offset=self.shard_offset,
                                   n_total=self.shard_total if self.process_group is not None else S, group=self.process_group)

    def orthonormalize_left(self):
        raise NotImplementedError("QR re-gauge of a conv-TT")

    orthonormalize_right = orthonormalize_left
    node_orthonormalize_left = node_orthonormalize_right = lambda self, node: self.orthonormalize_left()
