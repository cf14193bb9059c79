"""Sweep engine: per-site Gauss-Newton / ALS on B200.

Same public surface as the reference's ``TensorNetwork`` (tensor/network.py:13-932): a layer
hands over its node graph and callers drive ``accumulating_swipe`` / ``lanczos_swipe`` /
``scipy_swipe`` / ``forward`` / ``forward_batch`` / ``orthonormalize_*`` with the reference's
keywords and return conventions.  Underneath nothing is shared with
it: the graph is *recognised* once as a chain of sites, environments are cached per site
(O(sites) work per sweep instead of the reference's O(sites^2)), and all arithmetic runs in the
hand-written CUDA kernels of libtn_b200.so through ``ops``.  There is no CPU path.

Parameters stay ordinary ``torch.Tensor`` objects on ``node.tensor`` (a new tensor per update),
so ``node_states()/load_node_states()`` snapshots and EarlyStopping keep working; any outside
change of a core is noticed through tensor identity/version and drops the cached environments,
the way ``set_input``/``reset_stacks`` do in the reference (network.py:78-81, 329-345).
"""
import os
import time

import torch

from .. import ops
from ..ops import Factor
from .bregman import hessian_terms
from .node import TensorNode

_GRAM_MODES = {"fp64": ops.GRAM_FP64, "tf32": ops.GRAM_TF32, "tf32x3": ops.GRAM_TF32X3, "3xtf32": ops.GRAM_TF32X3, "f16": ops.GRAM_F16}


class MappedInput:
    """Raw samples ``X (N, F)`` plus a per-feature map evaluated inside the kernels.

    Stands in for the list of pre-mapped ``(N, f)`` tensors the reference builds with ``fbasis`` /
    ``polynomial_basis`` (models/tnml.py:11-23): site k reads column k of X and maps it on the fly,
    so the n*f*N*8 bytes of mapped copies never exist.
    """

    def __init__(self, X, kind="sin-cos", degree=3):
        if kind not in ("sin-cos", "polynomial"):
            raise ValueError(f"unknown feature map {kind!r}")
        self.X = X
        self.kind = kind
        self.f = 2 if kind == "sin-cos" else degree + 1
        self.map_kind = ops.MAP_SINCOS if kind == "sin-cos" else ops.MAP_POLY

    def __len__(self):
        return self.X.shape[0]

    @property
    def device(self):
        return self.X.device

    def __getitem__(self, sl):
        out = MappedInput.__new__(MappedInput)
        out.X, out.kind, out.f, out.map_kind = self.X[sl], self.kind, self.f, self.map_kind
        return out

    def to(self, *a, **k):
        out = MappedInput.__new__(MappedInput)
        out.X, out.kind, out.f, out.map_kind = self.X.to(*a, **k), self.kind, self.f, self.map_kind
        return out


class _Site:
    __slots__ = ("node", "left", "right", "phys", "cls", "input_index", "linear")


def sweep_schedule(first_cols, second_cols, num_swipes, eps, eps_decay=None, skip_second=False, direction="l2r",
                   eps_per_node=False):
    """Order, half-sweep counter NS and epsilon of every update ``accumulating_swipe`` performs.

    Pure function of the reference's control flow (network.py:409-433, 506-535, 604).
    ``first_cols`` / ``second_cols`` are the column ids (``node_indices``) of the nodes of the first and
    second half-sweep in visiting order.  Returns ``(NS, half, index_in_half, eps)`` tuples.  The
    turn-around skip compares columns: a node is skipped when the previous half ended on its column.
    """
    out = []
    NS = 0
    last_first = None
    last_second = None

    def pick(idx):
        # A list that is too short is an IndexError in the reference at the moment the sweep gets there -- and never, when a
        # convergence criterion ends the sweep first (TensorTrainRegressorEarlyStopping passes one epsilon per CORE while a
        # linear-projection train has two trained nodes per core, tensor/module.py:573-582): the error object takes the place of the
        # value and accumulating_swipe raises it when it reaches the entry.
        if not isinstance(eps, list):
            return eps
        try:
            return eps[idx]
        except IndexError as err:
            return err

    def eps_at(ns):
        e = pick(ns)
        if eps_decay is not None and not isinstance(e, Exception):
            e = e * eps_decay ** ns
        return e

    for _ in range(num_swipes):
        e = eps_at(NS)
        for i, col in enumerate(first_cols):
            if eps_per_node and not isinstance(e, Exception):
                e = pick(i if direction == "l2r" else len(first_cols) - 1 - i)
            last_first = col
            if last_second is not None and col == last_second:
                continue
            out.append((NS, 0, i, e))
        NS += 1
        if skip_second:
            continue
        e = eps_at(NS)
        for i, col in enumerate(second_cols):
            if eps_per_node and not isinstance(e, Exception):
                e = pick(i if direction == "r2l" else len(second_cols) - 1 - i)
            last_second = col
            if last_first is not None and col == last_first:
                continue
            out.append((NS, 1, i, e))
        NS += 1
    return out


def batch_mean_of_means(loss_rows, batch_size, row_offset=0, n_total=None, group=None):
    """mean over minibatches of the per-minibatch mean loss (reference network.py:472-474): the last,
    short batch weighs as much as a full one.  ``loss_rows`` is (S_local, ...) for global rows
    [row_offset, row_offset+S_local); with ``group`` the per-batch sums are all-reduced."""
    S = loss_rows.shape[0]
    per = loss_rows.reshape(S, -1).mean(dim=1) if loss_rows.dim() > 1 else loss_rows
    N = n_total if n_total is not None else S
    bs = N if batch_size <= 0 or batch_size > N else batch_size
    nb = (N + bs - 1) // bs
    if nb == 1 and group is None:
        return per.mean()
    idx = (torch.arange(S, device=per.device) + row_offset) // bs
    sums = torch.zeros(nb, dtype=per.dtype, device=per.device).index_add_(0, idx, per)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(sums, group=group)
    counts = torch.full((nb,), float(bs), dtype=per.dtype, device=per.device)
    counts[-1] = float(N - (nb - 1) * bs)
    return (sums / counts).mean()


class TensorNetwork:
    _supports_gradient = True       # method='gradient' (per-minibatch first-order steps); engines with their own update say False

    def __init__(self, input_nodes, main_nodes, train_nodes=None, output_labels=("s",), sample_dim="s"):
        self.input_nodes = input_nodes
        self.main_nodes = main_nodes
        self.train_nodes = main_nodes if train_nodes is None else train_nodes
        self.output_labels = output_labels
        self.sample_dim = sample_dim
        self.left_stacks = None   # kept for API compatibility; the engine uses its own caches
        self.right_stacks = None
        self.nodes, self.node_indices = self._discover_nodes()
        self.gram_mode = "fp64"
        # local solve: "fp64" = blocked fp64 Cholesky; "mixed" = tensor-core factorisation + fp64 refinement;
        # "auto" = mixed for large systems whose Gram was itself built on the tensor cores, fp64 otherwise
        self.solve_mode = "auto"
        self.mixed_min_P = 8192
        self.mixed_rtol = 1e-9          # residual the refinement aims for ...
        self.mixed_accept = 1e-9        # ... and the one it must reach for the result to be used (else fp64 redo)
        self.solve_stats = {"mixed": 0, "fp64": 0, "mixed_fallback": 0, "refine_iters": 0, "refined": 0, "gram_fp64_fallback": 0}
        self._mixed_floor = 0.0         # ridge values at or below this made the mixed solve fall back; skip it there
        # tensor-core Gram modes: "exact" = the TF32 / 3xTF32 Gram only preconditions (through its Cholesky factor) a conjugate-
        # gradient iteration on the fp64 matrix-free operator J^T W J / sigma + ridge, so the step solves the reference's fp64
        # system (network.py:293-327) to `refine_rtol`; "gram" = round 1's behaviour (the step solves the tensor-core Gram's system)
        self.refine = "exact"
        self.refine_rtol = 1e-11        # stopping test of the refinement: estimated relative forward error of the step ...
        self.refine_accept = 1e-9       # ... and what it must reach for the step to be used, else the site is redone with an fp64 Gram
        self.refine_max_iter = 30
        # tensor-core factorisation of the preconditioner: 1 = 3xTF32 trailing updates, 2 = one TF32 pass (TN_FACTOR_ONE_PASS=1)
        self.factor_passes_code = 2 if os.environ.get("TN_FACTOR_ONE_PASS", "0") not in ("", "0") else 1
        self._refine_floor = -1.0       # ridge values at or below this needed the fp64 Gram; go there directly
        # fp32 accumulation window (rows) of the tensor-core Gram when it only preconditions the exact refinement: longer = faster,
        # coarser (None = the library default of 2048, the window the stand-alone accuracy figures of the Gram are quoted for)
        self.tc_flush_rows = 16384
        self.small_site_fp64 = 2.0e9    # rows x unique Gram entries below which a site is built in fp64 outright (~0.5 ms of DMMA)
        self.process_group = None       # torch.distributed group: x, y are then this rank's row shard
        self.shard_offset = 0           # global index of this rank's first row
        self.shard_total = None         # global number of rows
        self.lean_envs = False
        self._sites = None
        self._left = {}
        self._right = {}
        self._stamps = None
        self._data_key = None
        self._data = None
        self.last_site_seconds = None
        self.on_data_ready = None       # optional hook, called once the (possibly host-resident) data has been enqueued to the device

    # ------------------------------------------------------------------ graph / plan
    def _discover_nodes(self):
        idx = {n: i for i, n in enumerate(self.main_nodes)}
        seen = list(self.main_nodes)
        queue = list(self.main_nodes)
        while queue:
            cur = queue.pop(0)
            for lab, other in cur.connections.items():
                if other not in idx and not cur.is_horizontal_bond(lab):
                    idx[other] = idx[cur]
                    seen.append(other)
                    queue.append(other)
        return sorted(seen, key=lambda n: n.name), idx

    def cuda(self):
        for n in self.nodes:
            n.cuda()
        return self

    def to(self, device=None, dtype=None):
        for n in self.nodes:
            n.to(device=device, dtype=dtype)
        return self

    def _plan(self):
        if self._sites is not None:
            return self._sites
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        sites = []
        for node in self.main_nodes:
            s = _Site()
            s.node = node
            s.left = node.left_labels[0] if node.left_labels else None
            s.right = node.right_labels[0] if node.right_labels else None
            if len(node.left_labels) > 1 or len(node.right_labels) > 1:
                raise NotImplementedError(f"{node.name}: more than one bond per side is not a plain chain")
            cls = [l for l in node.dim_labels if l in out_labels]
            if len(cls) > 1:
                raise NotImplementedError(f"{node.name}: more than one output leg")
            s.cls = cls[0] if cls else None
            phys = [(lab, other) for lab in node.dim_labels for l2, other in node.connections.items()
                    if l2 == lab and any(other is n for n in self.input_nodes)]
            s.linear = None
            if not phys:
                # TensorTrainLinearLayer (reference layers.py:308-343): core -[lin]- W (lin, p) -[p]- input
                for lab in node.dim_labels:
                    other = node.connections.get(lab)
                    if other is None or node.is_horizontal_bond(lab) or other.tensor.dim() != 2 or len(other.dim_labels) != 2:
                        continue
                    far = [(l2, n2) for l2, n2 in other.connections.items() if l2 != lab and any(n2 is n for n in self.input_nodes)]
                    if len(far) == 1 and other.dim_labels == [lab, far[0][0]]:
                        s.linear = other
                        phys = [(lab, far[0][1])]
                        break
            if not phys:
                phys = self._phys_behind_operator(node)
            if len(phys) not in (1, 2):
                raise NotImplementedError(f"{node.name}: expected one input node (or two for a 2-site block), found {len(phys)} "
                                          "(this engine covers tensor-train chains; see DESIGN.md)")
            # a 2-site (DMRG) block carries two physical legs; its site input is the Kronecker product of both
            s.phys = phys[0][0] if len(phys) == 1 else tuple(lab for lab, _ in phys)
            idxs = [next(i for i, n in enumerate(self.input_nodes) if n is other) for _, other in phys]
            s.input_index = idxs[0] if len(phys) == 1 else tuple(idxs)
            known = ({s.left, s.right, s.cls} | set(lab for lab, _ in phys)) - {None}
            extra = [l for l in node.dim_labels if l not in known]
            if extra:
                raise NotImplementedError(f"{node.name}: unsupported legs {extra}")
            sites.append(s)
        for a, b in zip(sites[:-1], sites[1:]):
            if a.right is None or a.right != b.left:
                raise NotImplementedError(f"{a.node.name}-{b.node.name}: bond labels do not chain")
        if len([s for s in sites if s.cls is not None]) > 1:
            raise NotImplementedError("more than one site carries an output leg")
        self._sites = sites
        return sites

    def _phys_behind_operator(self, node):
        """[(physical label, input node)] when an operator node sits between the core and its input (engines that replace such an
        operator by a closed form override this); none in a plain chain."""
        return []

    def _owner(self):
        for k, s in enumerate(self._plan()):
            if s.cls is not None and s.node.dim_size(s.cls) > 1:
                return k
        return None

    def _canon(self, k):
        """Core k as a (r_l, c, f, r_r) tensor (a view when the label order already is canonical)."""
        s = self._plan()[k]
        node = s.node
        order = self._order(s)
        t = node.tensor.permute(*[node.dim_labels.index(l) for l in order])
        rl = node.dim_size(s.left) if s.left in node.dim_labels else 1
        c = node.dim_size(s.cls) if s.cls in node.dim_labels else 1
        f = self._phys_size(s)
        rr = node.dim_size(s.right) if s.right in node.dim_labels else 1
        return t.reshape(rl, c, f, rr)

    @staticmethod
    def _phys_labels(s):
        return list(s.phys) if isinstance(s.phys, tuple) else [s.phys]

    def _phys_size(self, s):
        f = 1
        for lab in self._phys_labels(s):
            f *= s.node.dim_size(lab)
        return f

    def _order(self, s):
        """Canonical leg order (left, class, physical..., right) restricted to the legs the node has."""
        cand = [s.left, s.cls] + self._phys_labels(s) + [s.right]
        return [l for l in cand if l is not None and l in s.node.dim_labels]

    def _from_canon(self, k, t4):
        """(r_l, c, f, r_r) tensor -> tensor in the node's own label order and shape."""
        s = self._plan()[k]
        node = s.node
        order = self._order(s)
        # bond and class sizes come from t4 (a QR re-gauge of a wide core shrinks a bond, reference network.py:644-657), the
        # physical legs keep the node's
        size = {s.left: t4.shape[0], s.cls: t4.shape[1], s.right: t4.shape[3]}
        t = t4.reshape([size[l] if l in size else node.dim_size(l) for l in order])
        return t.permute(*[order.index(l) for l in node.dim_labels]).contiguous()

    # ------------------------------------------------------------------ data binding / caches
    def _bind(self, x):
        """Per-site Factor descriptors for a data object (tensor shared by all sites, list, MappedInput)."""
        sites = self._plan()
        facs = []
        if isinstance(x, MappedInput):
            X = x.X
            for k, s in enumerate(sites):
                if isinstance(s.phys, tuple):
                    raise NotImplementedError("2-site blocks take explicit inputs, not a fused feature map")
                if s.node.dim_size(s.phys) != x.f:
                    raise ValueError(f"site {k}: core has f={s.node.dim_size(s.phys)}, feature map gives {x.f}")
                facs.append(Factor(X, m=x.f, map_kind=x.map_kind, col=s.input_index))
            return facs, X.shape[0], X.device

        def tensor_of(j):
            t = x[j] if isinstance(x, (list, tuple)) else x
            if t.dim() != 2 or t.stride(1) != 1:
                t = t.reshape(t.shape[0], -1).contiguous()
            return t

        for k, s in enumerate(sites):
            if isinstance(s.phys, tuple):
                ta, tb = tensor_of(s.input_index[0]), tensor_of(s.input_index[1])
                fa, fb = (s.node.dim_size(l) for l in s.phys)
                if ta.shape[1] != fa or tb.shape[1] != fb:
                    raise ValueError(f"site {k}: inputs have {ta.shape[1]}x{tb.shape[1]} features, block expects {fa}x{fb}")
                t = (ta[:, :, None] * tb[:, None, :]).reshape(ta.shape[0], fa * fb).contiguous()   # phi_L (x) phi_R per sample
            else:
                t = tensor_of(s.input_index)
                if s.linear is None and t.shape[1] != s.node.dim_size(s.phys):
                    raise ValueError(f"site {k}: input has {t.shape[1]} features, core expects {s.node.dim_size(s.phys)}")
            facs.append(Factor(t, m=t.shape[1]))
        t0 = tensor_of(0)
        return facs, t0.shape[0], t0.device

    def _project(self, facs, S):
        """Site inputs after the linear projections (phi_k = x W_k^T for sites with a linear node); returns (facs, raw)."""
        sites = self._plan()
        if not any(st.linear is not None for st in sites):
            return facs, None
        out = []
        for st, fac in zip(sites, facs):
            if st.linear is None:
                out.append(fac)
                continue
            W = st.linear.tensor                                   # (lin, p)
            if fac.m != W.shape[1]:
                raise ValueError(f"{st.linear.name}: input has {fac.m} features, projection expects {W.shape[1]}")
            phi = ops.env_update(None, fac, W.t().reshape(1, W.shape[1], W.shape[0]), S)
            out.append(Factor(phi, m=W.shape[0]))
        return out, facs

    @staticmethod
    def _key_of(x):
        if isinstance(x, MappedInput):
            return ("map", id(x.X), x.kind, x.f)
        if isinstance(x, (list, tuple)):
            return ("list",) + tuple(id(t) for t in x)
        return ("tensor", id(x))

    def set_input(self, x):
        """Bind training data; a different object drops the cached environments (network.py:329-345)."""
        key = self._key_of(x)
        if key == self._data_key:
            return False
        self._data_key = key
        self._rebind(x)
        return True

    def _rebind(self, x):
        facs, S, dev = self._bind(x)
        if any(st.linear is not None for st in self._plan()):
            self._require_cuda(dev)
        facs, self._raw_facs = self._project(facs, S)
        self._data = (x, facs, S, dev)      # keep x alive so ids stay unique
        self._left.clear()
        self._right.clear()

    def reset_stacks(self, node=None):
        self._left.clear()
        self._right.clear()
        self.left_stacks = None
        self.right_stacks = None

    # The reference keeps its environments as dictionaries of contracted nodes and exposes their maintenance (network.py:73-99,152-172);
    # callers outside the class only ever ask for them to be brought up to date (symmetric_operator.py:52, cum_sum_operator.py:67).
    # Here the environments are cached device tensors rebuilt on demand, so "recompute" and "update" mean: forget what the change
    # invalidated.
    def recompute_all_stacks(self, exclude_nodes=None):
        self.reset_stacks()

    def left_update_stacks(self, node):
        """The node changed: environments that contain it are dropped (reference network.py:152-161 recontracts the left one)."""
        if type(self) is TensorNetwork and any(node is n for n in self.main_nodes):
            self._core_changed(self.main_nodes.index(node))
        else:
            self.reset_stacks()          # subclasses keep further caches: drop everything

    right_update_stacks = left_update_stacks

    def _stamp(self):
        return [(id(s.node.tensor), s.node.tensor._version) + ((id(s.linear.tensor), s.linear.tensor._version) if s.linear is not None else ())
                for s in self._plan()]

    def _check_external(self):
        st = self._stamp()
        if self._stamps != st:
            if self._data is not None and getattr(self, "_raw_facs", None) is not None:
                self._rebind(self._data[0])      # a projection changed behind our back: re-project the bound data
            self._left.clear()
            self._right.clear()
            self._stamps = st

    def _core_changed(self, k):
        for j in [j for j in self._left if j >= k]:
            del self._left[j]
        for j in [j for j in self._right if j <= k]:
            del self._right[j]
        self._stamps = self._stamp()

    # ------------------------------------------------------------------ environments
    def _step(self, env, fac, k, left, rows_S):
        """One environment step through site k.  env: (S, c, r) or None; returns (S, c', r')."""
        G = self._canon(k)
        rl, c, f, rr = G.shape
        S = rows_S
        if left:
            r_in, r_out = rl, rr
            core = G if c == 1 else G.permute(0, 2, 1, 3)          # (rl, f, c, rr)
        else:
            r_in, r_out = rr, rl
            core = G.permute(3, 2, 1, 0)                           # (rr, f, c, rl)
        cin = 1 if env is None else env.shape[1]
        if c == 1:
            core3 = (core.reshape(rl, f, rr) if left else core.reshape(rr, f, rl))
            rows = S * cin
            e2 = None if env is None else env.reshape(rows, r_in)
            out = ops.env_update(e2, fac, core3, rows, cdiv=cin)
            return out.view(S, cin, r_out)
        if cin != 1:
            raise NotImplementedError("two class legs meet")
        core3 = core.reshape(r_in, f, c * r_out)
        e2 = None if env is None else env.reshape(S, r_in)
        out = ops.env_update(e2, fac, core3, S, cdiv=1)
        return out.view(S, c, r_out)

    def _get_left(self, k):
        """Environment of sites 0..k (None for k < 0)."""
        if k < 0:
            return None
        _, facs, S, _ = self._data
        j = k
        while j >= 0 and j not in self._left:
            j -= 1
        env = self._left[j] if j >= 0 else None
        for i in range(j + 1, k + 1):
            env = self._step(env, facs[i], i, True, S)
            self._left[i] = env
        return env

    def _get_right(self, k):
        """Environment of sites k..n-1 (None for k >= n)."""
        n = len(self._plan())
        if k >= n:
            return None
        _, facs, S, _ = self._data
        j = k
        while j < n and j not in self._right:
            j += 1
        env = self._right[j] if j < n else None
        for i in range(j - 1, k - 1, -1):
            env = self._step(env, facs[i], i, False, S)
            self._right[i] = env
        return env

    # ------------------------------------------------------------------ forward
    def _chain_forward(self, x):
        """Prediction (S, C) for arbitrary data, without touching the training caches."""
        facs, S, dev = self._bind(x)
        self._require_cuda(dev)
        facs, _ = self._project(facs, S)
        env = None
        for k in range(len(self._plan())):
            env = self._step(env, facs[k], k, True, S)
        return env[:, :, 0]

    def _require_cuda(self, dev):
        if dev.type != "cuda":
            raise RuntimeError("tensornetworksfork_b200 runs on CUDA devices only (no CPU fallback); move data and "
                               "model with .to('cuda')")
        for n in self.main_nodes:
            if n.tensor.device != dev:
                raise RuntimeError(f"core {n.name} lives on {n.tensor.device}, data on {dev}")

    def forward(self, x, to_tensor=False):
        """Prediction.  Returns a TensorNode labelled by ``output_labels`` (reference network.py:115-137)
        or the bare tensor."""
        y = self._chain_forward(x)
        C = self._num_outputs()
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        if not out_labels:
            y = y[:, 0]
        if to_tensor:
            return y
        return TensorNode(y, [self.sample_dim] + out_labels, name="O")

    def forward_batch(self, x, batch_size):
        """Reference network.py:139-150.  Batching only bounds temporary memory here."""
        n = len(x) if not isinstance(x, (list, tuple)) else x[0].shape[0]
        if batch_size <= 0 or batch_size >= n:
            return self.forward(x, to_tensor=True)
        chunk = max(batch_size, 1 << 16)  # same numbers, fewer launches: rows are independent
        outs = []
        for lo in range(0, n, chunk):
            xb = x[lo:lo + chunk] if not isinstance(x, (list, tuple)) else [t[lo:lo + chunk] for t in x]
            outs.append(self.forward(xb, to_tensor=True))
        return torch.cat(outs, dim=0)

    def _num_outputs(self):
        k = self._owner()
        if k is None:
            return 1
        s = self._plan()[k]
        return s.node.dim_size(s.cls)

    # ------------------------------------------------------------------ local system of one site
    def _site_problem(self, k, y, loss_fn, rows=None):
        """Prediction, loss terms and the three Kronecker factors of site k's Jacobian.

        Returns dict with: yhat (S,C), loss (S[,C]), gram factors + weights + rows, rhs factors + weights,
        m_pos (sizes of the three parameter positions in canonical (a,c,p,b) order, merged to three).
        ``rows = (lo, hi)`` restricts the problem to that range of the bound rows (one minibatch of ``method='gradient'``).
        """
        _, facs, S, dev = self._data
        G = self._canon(k)
        rl, ck, f, rr = G.shape
        L = self._get_left(k - 1)
        R = self._get_right(k + 1)
        owner = self._owner()
        C = self._num_outputs()
        xk = facs[k]
        yoff = getattr(self, "_yhat_offset", None)
        if rows is not None:
            lo, hi = rows
            L = None if L is None else L[lo:hi]
            R = None if R is None else R[lo:hi]
            xk = Factor(xk.tensor[lo:hi], m=xk.m, div=xk.div, map_kind=xk.map_kind, col=xk.col)
            y = y[lo:hi]
            yoff = None if yoff is None else yoff[lo:hi]
            S = hi - lo
        if self.gram_mode != "fp64" and xk.map_kind != ops.MAP_IDENTITY:
            # tensor-core Gram: evaluate the feature map once per site (S x f) so the kernel stays on its fast path
            phi = ops.env_update(None, xk, torch.eye(f, dtype=torch.float64, device=dev).reshape(1, f, f), S)
            xk = Factor(phi, m=f)
        one = ops.ones_factor(G)

        def fac_of(env, r, div):
            if env is None:
                return one
            return Factor(env.reshape(-1, r), m=r, div=div)

        # ---- prediction
        if C == 1:
            Lf = None if L is None else L.reshape(S, rl)
            Rf = R.reshape(S, rr) if R is not None else torch.ones((1, 1), dtype=torch.float64, device=dev)
            yhat = ops.predict(Lf, xk, G.reshape(rl, f, rr), Rf, S, dot_div=1 if R is not None else (1 << 30)).view(S, 1)
        elif owner < k:
            Rf = R.reshape(S, rr) if R is not None else torch.ones((1, 1), dtype=torch.float64, device=dev)
            yhat = ops.predict(L.reshape(S * C, rl), xk, G.reshape(rl, f, rr), Rf, S * C, cdiv=C,
                               dot_div=C if R is not None else (1 << 30)).view(S, C)
        elif owner > k:
            Lf = None if L is None else L.reshape(S, rl)
            yhat = ops.predict(Lf, xk, G.reshape(rl, f, rr), R.reshape(S * C, rr), S * C, cdiv=C, env_div=C).view(S, C)
        else:
            Lf = None if L is None else L.reshape(S, rl)
            Rf = R.reshape(S, rr) if R is not None else torch.ones((1, 1), dtype=torch.float64, device=dev)
            yT = torch.empty((C, S), dtype=torch.float64, device=dev)
            for c in range(C):
                ops.predict(Lf, xk, G[:, c].contiguous(), Rf, S, dot_div=1 if R is not None else (1 << 30), out=yT[c])
            yhat = yT.t().contiguous()

        if yoff is not None:
            yhat = yhat + yoff                       # outputs of the other members of a SumOfNetworks (held fixed)
        out_labels = [l for l in self.output_labels if l != self.sample_dim]
        y_in = yhat if out_labels else yhat[:, 0]
        loss, g, U, lam = hessian_terms(loss_fn, y_in, y)
        g = g.reshape(S, C).contiguous()
        V = lam.shape[1]

        # ---- factors
        if C == 1:
            w = (lam.reshape(S, V) * U.reshape(S, V) ** 2).sum(dim=1).contiguous()
            fa, fb, fc = fac_of(L, rl, 1), xk, fac_of(R, rr, 1)
            prob = dict(gram=(fa, fb, fc), gw=w, grows=S, rhs=(fa, fb, fc), rw=g.reshape(S).contiguous(), rrows=S,
                        m_pos=(rl, f, rr))
        else:
            U = U.contiguous()
            lamf = lam.reshape(S * V).contiguous()
            xv = Factor(xk.tensor, m=xk.m, div=V, map_kind=xk.map_kind, col=xk.col)
            if owner < k:
                F, Gr = ops.class_rows(L, U, g)
                prob = dict(gram=(Factor(F, m=rl), xv, fac_of(R, rr, V)), gw=lamf, grows=S * V,
                            rhs=(Factor(Gr, m=rl), xk, fac_of(R, rr, 1)), rw=None, rrows=S, m_pos=(rl, f, rr))
            elif owner > k:
                F, Gr = ops.class_rows(R, U, g)
                prob = dict(gram=(fac_of(L, rl, V), xv, Factor(F, m=rr)), gw=lamf, grows=S * V,
                            rhs=(fac_of(L, rl, 1), xk, Factor(Gr, m=rr)), rw=None, rrows=S, m_pos=(rl, f, rr))
            else:
                if rl == 1:
                    Fu = U.reshape(S * V, C)
                    Fg = g
                else:  # class leg on an interior site: merge (a, c) into one factor
                    Lf = L.reshape(S, rl)
                    Fu = (Lf[:, None, :, None] * U[:, :, None, :]).reshape(S * V, rl * C).contiguous()
                    Fg = (Lf[:, :, None] * g[:, None, :]).reshape(S, rl * C).contiguous()
                prob = dict(gram=(Factor(Fu, m=rl * C), xv, fac_of(R, rr, V)), gw=lamf, grows=S * V,
                            rhs=(Factor(Fg, m=rl * C), xk, fac_of(R, rr, 1)), rw=None, rrows=S, m_pos=(rl * C, f, rr))
        prob["yhat"] = yhat
        prob["loss"] = loss
        prob["keep"] = (L, R, U, lam, g)  # keep operands alive until the kernels have been enqueued
        return prob

    @staticmethod
    def _roles(m_pos):
        """Which parameter position becomes the N dimension of the Gram GEMM (least padded work)."""
        n = [ops.npairs(m) for m in m_pos]
        best = None
        for c in range(3):
            others = [t for t in range(3) if t != c]
            nu = n[others[0]] * n[others[1]]
            cost = (-(-nu // 128) * 128) * (-(-n[c] // 64) * 64)
            if best is None or cost < best[0]:
                best = (cost, c, others)
        _, c, others = best
        role_of_pos = [0, 0, 0]
        role_of_pos[others[0]] = 0
        role_of_pos[others[1]] = 1
        role_of_pos[c] = 2
        return role_of_pos, (others[0], others[1], c)

    def _accumulate(self, prob, gram_mode=None):
        """Local Gram (unique entries M), right-hand side b and the bookkeeping needed to expand them.  In the tensor-core
        modes the exact fp64 trace of the Gram rides along (two numbers) for the scaling of the system."""
        gram_mode = self.gram_mode if gram_mode is None else gram_mode
        mode = _GRAM_MODES[gram_mode]
        m_pos = prob["m_pos"]
        role_of_pos, order = self._roles(m_pos)
        gf = prob["gram"]
        nM = ops.npairs(m_pos[0]) * ops.npairs(m_pos[1]) * ops.npairs(m_pos[2])
        P = m_pos[0] * m_pos[1] * m_pos[2]
        dev = prob["yhat"].device
        extra = 2 if gram_mode != "fp64" else 0
        buf = torch.empty((nM + P + extra,), dtype=torch.float64, device=dev)
        M, b = buf[:nM], buf[nM:nM + P]
        flush = self.tc_flush_rows if (extra and self.refine == "exact") else None
        ops.gram(mode, gf[order[0]], gf[order[1]], gf[order[2]], prob["gw"], prob["grows"], M=M, flush_rows=flush)
        rf = prob["rhs"]
        ops.rhs(rf[0], rf[1], rf[2], prob["rw"], prob["rrows"], b=b)
        if extra:
            ops.gram_trace(gf[0], gf[1], gf[2], prob["gw"], prob["grows"], out=buf[nM + P:])
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(buf, group=self.process_group)
        prob["trace"] = buf[nM + P:] if extra else None
        prob["gram_mode_used"] = gram_mode
        return M, b, role_of_pos

    def _solve(self, k, M, b, m_pos, role_of_pos, method, eps, prob=None):
        """sigma-scaling, ridge and Cholesky solve (reference network.py:293-327) -> step in node layout."""
        step = self._solve_flat(self._canon(k).contiguous().view(-1), M, b, m_pos, role_of_pos, method, eps, prob=prob)
        return self._from_canon(k, step.reshape(self._canon(k).shape))

    @staticmethod
    def _ridge_of(method, eps):
        m = method.lower()
        return 0.0 if m in ("exact", "cholesky", "gradient") or m.startswith("gradient") else 2.0 * float(eps)

    def _gram_mode_for(self, method, eps, prob=None):
        """Gram mode of one site update: the tensor-core mode, unless a ridge this small already needed the fp64 Gram, or the
        site is so small that the fp64 Gram costs less than the refinement's factor-and-iterate overhead (config 1: 4177 rows,
        P <= 324 -- a few launches of latency either way)."""
        if self.gram_mode == "fp64" or self.refine != "exact":
            return self.gram_mode
        if self._ridge_of(method, eps) <= self._refine_floor:
            return "fp64"
        if prob is not None and self.small_site_fp64 > 0:
            m = prob["m_pos"]
            work = float(prob["grows"]) * ops.npairs(m[0]) * ops.npairs(m[1]) * ops.npairs(m[2])
            if self.process_group is None and work <= self.small_site_fp64:
                return "fp64"
        return self.gram_mode

    def _solve_refined(self, theta, M, b, m_pos, role_of_pos, ridge, prob):
        """Tensor-core Gram modes: Cholesky factor of the TF32 / 3xTF32 Gram as the preconditioner of conjugate gradients on the
        fp64 matrix-free operator of the same rows.  The iteration solves  (J^T W J / sigma + ridge) x = -(b / sigma + ridge theta)
        with sigma the exact mean diagonal (fp64 trace), i.e. the reference's own system, to ``refine_rtol``."""
        P = m_pos[0] * m_pos[1] * m_pos[2]
        sigma = ops.gram_sigma(M, m_pos, role_of_pos)
        tr = prob.get("trace")
        if tr is not None:
            # mean |A_ii| = trace / P whenever no diagonal entry is negative -- every positive semi-definite output Hessian, whatever
            # the signs of the individual virtual-row weights.  The exact trace is used when it agrees with the mean |diagonal| of the
            # tensor-core Gram to that Gram's accuracy; an indefinite Hessian with negative diagonal entries keeps the latter.
            exact = tr[0:1] / float(P)
            sigma = torch.where((exact - sigma).abs() <= 1e-3 * sigma.abs(), exact, sigma)
        A = ops.gram_expand(M, m_pos, role_of_pos, sigma, ridge)
        rhs = ops.rhs_prepare(b, theta, sigma, ridge)
        tensor_core = self.solve_mode != "fp64" and P >= self.mixed_min_P
        work, info = ops.cholesky_factor(A, tensor_core=(self.factor_passes_code if tensor_core else 0))
        op = ops.Operator(P, factors=prob["gram"], w=prob["gw"], rows=prob["grows"], group=self.process_group, sigma=sigma, ridge=ridge)
        x, stats = ops.cg(op, rhs, precond=(A, work, info), max_iter=self.refine_max_iter, rtol=self.refine_rtol)
        bad = int(info.item())
        st = stats.tolist()
        rel, iters = st[4], st[1]            # value of the stopping criterion (forward-error estimate), iterations
        if self.process_group is not None:
            # the ranks' iterates agree to rounding only (atomics in the reductions): rank 0's step and verdict are the ones applied
            import torch.distributed as dist
            flag = torch.tensor([float(bad), rel], dtype=torch.float64, device=x.device)
            src = dist.get_global_rank(self.process_group, 0)
            dist.broadcast(flag, src=src, group=self.process_group)
            dist.broadcast(x, src=src, group=self.process_group)
            bad, rel = int(flag[0].item()), float(flag[1].item())
        if bad != 0 or not (rel <= self.refine_accept):
            self._refine_floor = max(self._refine_floor, ridge)
            raise _NeedExactGram(f"refinement reached {rel:.2e} in {int(iters)} iterations (info {bad}) at ridge {ridge:g}")
        self.solve_stats["refined"] += 1
        self.solve_stats["refine_iters"] += int(iters)
        self.solve_stats["refine_max_rel"] = max(self.solve_stats.get("refine_max_rel", 0.0), float(rel))
        self.solve_stats["mixed" if tensor_core else "fp64"] += 1
        return x

    def _solve_flat(self, theta, M, b, m_pos, role_of_pos, method, eps, prob=None):
        """The local solve on flat canonical-order vectors; returns the flat step."""
        m = method.lower()
        if m == "gradient":
            return -b
        if m not in ("exact", "ridge_exact", "cholesky") and not m.startswith("ridge_cholesky"):
            raise ValueError(f"Unknown method: {method}")
        ridge = 0.0 if m in ("exact", "cholesky") else 2.0 * float(eps)
        if prob is not None and prob.get("gram_mode_used", "fp64") != "fp64" and self.refine == "exact":
            return self._solve_refined(theta, M, b, m_pos, role_of_pos, ridge, prob)
        sigma = ops.gram_sigma(M, m_pos, role_of_pos)
        A = ops.gram_expand(M, m_pos, role_of_pos, sigma, ridge)
        rhs = ops.rhs_prepare(b, theta, sigma, ridge)
        P = rhs.numel()
        use_mixed = self.solve_mode == "mixed" or (self.solve_mode == "auto" and self.gram_mode != "fp64"
                                                   and P >= self.mixed_min_P and ridge > self._mixed_floor)
        info = None
        if use_mixed:
            rhs0 = rhs.clone()
            info, stats = ops.cholesky_solve_mixed(A, rhs, rtol=self.mixed_rtol)
            bad, (rel, iters) = int(info.item()), stats.tolist()
            if bad == 0 and rel <= self.mixed_accept:
                self.solve_stats["mixed"] += 1
                self.solve_stats["refine_iters"] += int(iters)
            else:
                # the ~1e-5 factor lost positive definiteness or preconditions too weakly: redo in fp64 (the lower
                # triangle of A was overwritten, so expand again)
                self.solve_stats["mixed_fallback"] += 1
                self._mixed_floor = max(self._mixed_floor, ridge)
                A = ops.gram_expand(M, m_pos, role_of_pos, sigma, ridge, A=A)
                rhs = rhs0
                info = None
        lu_methods = m in ("exact", "ridge_exact")
        rhs_lu = None
        if info is None:
            self.solve_stats["fp64"] += 1
            rhs_lu = rhs.clone() if lu_methods else None
            info = ops.cholesky_solve(A, rhs)
        bad = int(info.item())
        if bad != 0 and lu_methods and rhs_lu is not None:
            # 'exact' / 'ridge_exact' are LU solves in the reference (torch.linalg.solve, network.py:305-310), which also accept
            # systems that are not numerically positive definite.  The Cholesky path is the fast one; when it reports a lost
            # pivot the system is expanded again and handed to the LU solver of the library (rare, off the hot path).  An exactly
            # singular matrix still raises LinAlgError there, as in the reference.
            A = ops.gram_expand(M, m_pos, role_of_pos, sigma, ridge, A=A)
            rhs = torch.linalg.solve(A[:, :P], rhs_lu)
            self.solve_stats["lu_fallback"] = self.solve_stats.get("lu_fallback", 0) + 1
            bad = 0
        if self.process_group is not None:
            # every rank solved the same all-reduced system, but the substitution / refinement kernels sum with atomics,
            # so the last bits may differ between ranks: rank 0's step is the one everybody applies (cores stay identical)
            import torch.distributed as dist
            flag = torch.tensor([float(bad)], dtype=torch.float64, device=rhs.device)
            src = dist.get_global_rank(self.process_group, 0)
            dist.broadcast(flag, src=src, group=self.process_group)
            dist.broadcast(rhs, src=src, group=self.process_group)
            bad = int(flag.item())
        if bad != 0:
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not positive-definite "
                f"(the leading minor of order {bad} is not positive-definite).")
        return rhs

    def get_A_b(self, node, grad=None, hessian=None, method=None, y=None, loss_fn=None):
        """Dense (A, b) of one node in the reference's layout (network.py:174-217), built from the
        currently bound data.  Provided for inspection/tests; the sweep never forms the dense A."""
        k = self.main_nodes.index(node)
        if loss_fn is None:
            loss_fn = _FixedTerms(grad, hessian)
        prob = self._site_problem(k, y, loss_fn)
        M, b, role_of_pos = self._accumulate(prob)
        one = torch.ones((1,), dtype=torch.float64, device=M.device)
        A = ops.gram_expand(M, prob["m_pos"], role_of_pos, one, 0.0)
        P = A.shape[0]
        dims, perm = self._layout(k)
        A = A[:, :P].reshape(dims + dims)
        A = A.permute(*(perm + [len(dims) + p for p in perm])).contiguous()
        return A, self._from_canon(k, b.reshape(tuple(self._canon(k).shape)))

    def get_b(self, node, grad):
        """b = J^T grad of one node, in the node's own shape (reference network.py:259-291: the right-hand side the matrix-free
        sweeps start from), built from the currently bound data by the right-hand-side kernel; J is not formed.  As for ``get_A_b``
        the data is what ``set_input`` bound last (the engine's ``forward`` does not rebind, so that predicting on validation rows
        inside a sweep leaves the cached environments of the training rows alone)."""
        if self._data is None:
            raise RuntimeError("get_b: no data bound -- call set_input(x) first")
        k = self.main_nodes.index(node)
        S = grad.shape[0]
        g = grad.reshape(S, -1)
        prob = self._site_problem(k, None, _FixedTerms(g, torch.zeros((S, g.shape[1], 1), dtype=g.dtype, device=g.device)))
        rf = prob["rhs"]
        b = ops.rhs(rf[0], rf[1], rf[2], prob["rw"], prob["rrows"])
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(b, group=self.process_group)
        return self._from_canon(k, b.reshape(tuple(self._canon(k).shape)))

    def _layout(self, k):
        """(sizes of the node's legs in canonical order, permutation canonical -> node label order)."""
        s = self._plan()[k]
        node = s.node
        order = self._order(s)
        return [node.dim_size(l) for l in order], [order.index(l) for l in node.dim_labels]

    def solve_system(self, node, A, b, method="exact", eps=0.0):
        """Dense-input variant of the local solve for API compatibility (network.py:293-327)."""
        P = b.numel()
        A_f = A.reshape(P, P)
        scale = A_f.diagonal().abs().mean()
        scale = torch.where(scale == 0, torch.ones_like(scale), scale)
        m = method.lower()
        if m == "gradient":
            return -b
        if m not in ("exact", "ridge_exact", "cholesky") and not m.startswith("ridge_cholesky"):
            raise ValueError(f"Unknown method: {method}")
        ridge = 0.0 if m in ("exact", "cholesky") else 2.0 * float(eps)
        lda = (P + 7) // 8 * 8
        Ap = torch.zeros((P, lda), dtype=torch.float64, device=A.device)
        Ap[:, :P] = A_f / scale
        Ap.diagonal().add_(ridge)
        rhs = -(b.reshape(P) / scale + ridge * node.tensor.reshape(P))
        rhs_lu = rhs.clone() if m in ("exact", "ridge_exact") else None
        info = ops.cholesky_solve(Ap, rhs)
        if int(info.item()) != 0:
            if rhs_lu is None:
                raise torch.linalg.LinAlgError("linalg.cholesky: The factorization could not be completed because the "
                                               "input is not positive-definite")
            # LU methods of the reference (network.py:305-310): library LU when the fast Cholesky path loses a pivot
            A_lu = A_f / scale
            A_lu.diagonal().add_(ridge)
            rhs = torch.linalg.solve(A_lu, rhs_lu)
        return rhs.reshape(b.shape)

    # ------------------------------------------------------------------ QR re-gauge
    def orthonormalize_left(self):
        for n in self.main_nodes:
            self.node_orthonormalize_left(n)

    def orthonormalize_right(self):
        for n in self.main_nodes:
            self.node_orthonormalize_right(n)

    @staticmethod
    def _qr_reduced(a):
        """Reduced QR of ``a`` (m x n, contiguous; overwritten when tall): (Q (m x k), R (k x n)), k = min(m, n), in LAPACK's sign
        convention like ``torch.linalg.qr(mode='reduced')`` (reference network.py:644,686).  A wide matrix -- a core with fewer rows
        than columns, e.g. the first core of an unconstricted train with r > f -- has Householder vectors that only involve its first
        m columns: Q is the Q of that square block, R = [R1 | Q^T A2]."""
        m, n = a.shape
        if m >= n:
            R = ops.qr(a)
            return a, R
        q = a[:, :m].contiguous()
        R1 = ops.qr(q)
        rest = a[:, m:].contiguous()
        R2 = ops.env_update(None, Factor(q.t().contiguous(), m=m), rest.reshape(1, m, n - m), m)
        return q, torch.cat([R1, R2.reshape(m, n - m)], dim=1).contiguous()

    def node_orthonormalize_left(self, node):
        """core_k <- Q, core_{k+1} <- R core_{k+1} (reference network.py:625-660); the bond shrinks to the row count of a wide core."""
        k = self.main_nodes.index(node)
        if k >= len(self.main_nodes) - 1:
            return
        self._require_cuda(node.tensor.device)
        G = self._canon(k)
        rl, c, f, rr = G.shape
        a = G.reshape(rl * c * f, rr).contiguous().clone()
        Q, Rm = self._qr_reduced(a)                                   # (rows, kk), (kk, rr)
        kk = Q.shape[1]
        node.tensor = self._from_canon(k, Q.reshape(rl, c, f, kk))
        Gn = self._canon(k + 1)
        nl, nc, nf, nr = Gn.shape
        newn = ops.env_update(None, Factor(Rm, m=rr), Gn.reshape(1, nl, nc * nf * nr), kk)
        self.main_nodes[k + 1].tensor = self._from_canon(k + 1, newn.reshape(kk, nc, nf, nr))
        self._core_changed(k)
        self._core_changed(k + 1)

    def node_orthonormalize_right(self, node):
        """RQ via the doubly flipped QR; factor pushed into core_{k-1} (reference network.py:662-707)."""
        k = self.main_nodes.index(node)
        if k <= 0:
            return
        self._require_cuda(node.tensor.device)
        G = self._canon(k)
        rl, c, f, rr = G.shape
        a = G.permute(1, 2, 3, 0).reshape(c * f * rr, rl)
        a = torch.flip(a, dims=[0, 1]).contiguous()
        Qrev, Rrev = self._qr_reduced(a)                              # (rows, kk), (kk, rl)
        kk = Qrev.shape[1]
        Q = torch.flip(Qrev, dims=[0, 1]).reshape(c, f, rr, kk).permute(3, 0, 1, 2)
        Rm = torch.flip(Rrev.t(), dims=[0, 1]).contiguous()          # (rl_old, rl_new)
        node.tensor = self._from_canon(k, Q.reshape(kk, c, f, rr))
        Gp = self._canon(k - 1)
        pl, pc, pf, pr = Gp.shape
        rows = pl * pc * pf
        newp = ops.env_update(None, Factor(Gp.reshape(rows, pr).contiguous(), m=pr), Rm.reshape(1, pr, kk), rows)
        self.main_nodes[k - 1].tensor = self._from_canon(k - 1, newp.reshape(pl, pc, pf, kk))
        self._core_changed(k)
        self._core_changed(k - 1)

    # ------------------------------------------------------------------ the sweep
    def _prepare_data(self, x, y_true, data_device, model_device):
        """Bind (and if needed move) the training data.  Host-resident data is copied once per call,
        not once per minibatch per site as the reference does (network.py:397-403,453-454)."""
        target = model_device if model_device is not None else self.main_nodes[0].tensor.device
        target = torch.device(target)

        def mv(t):
            if isinstance(t, MappedInput):
                return t if t.device == target else t.to(target, non_blocking=True)
            if isinstance(t, (list, tuple)):
                moved = [mv(u) for u in t]
                return t if all(a is b for a, b in zip(moved, t)) else moved
            return t if t.device == target else t.to(target, non_blocking=True)

        xm = mv(x)
        ym = mv(y_true)
        self._require_cuda(target)
        self.set_input(xm)
        if ym.dtype != torch.float64:
            ym = ym.to(torch.float64)
        if self.on_data_ready is not None:
            self.on_data_ready()
        return xm, ym

    def _one_update(self, k, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss):
        self._check_external()
        prob = self._site_problem(k, y, loss_fn)
        M, b, role_of_pos = self._accumulate(prob, self._gram_mode_for(method, eps, prob))
        try:
            step = self._solve(k, M, b, prob["m_pos"], role_of_pos, method, eps, prob=prob)
        except _NeedExactGram:
            # the tensor-core Gram was too coarse a preconditioner for this ridge (or lost positive definiteness): the site is
            # redone with the fp64 Gram and the fp64 factorisation -- loudly counted, and remembered for smaller ridges
            self.solve_stats["gram_fp64_fallback"] += 1
            del M, b
            M, b, role_of_pos = self._accumulate(prob, "fp64")
            step = self._solve(k, M, b, prob["m_pos"], role_of_pos, method, eps, prob=prob)
        node = self.main_nodes[k]
        new = node.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
        node.tensor = new
        self._core_changed(k)
        if not need_loss:
            return None
        S = prob["yhat"].shape[0]
        return batch_mean_of_means(prob["loss"], batch_size, row_offset=self.shard_offset,
                                   n_total=self.shard_total if self.process_group is not None else S,
                                   group=self.process_group)

    def _linear_site(self, node):
        """Index of the site whose linear-projection node `node` is, or None."""
        for k, st in enumerate(self._plan()):
            if getattr(st, "linear", None) is node:
                return k
        return None

    def _linear_problem(self, k, y, loss_fn):
        """Local problem of the projection W_k (lin, p) of a TensorTrainLinearLayer site (reference layers.py:308-343).

        J[s, c, (l, p)] = Q[s, c, l] x[s, p] with Q = L G_k R contracted over both bonds: a Kronecker product of two factors, so the
        Gram / rhs / matvec kernels of a core apply with m_pos = (lin, p, 1)."""
        prob = self._site_problem(k, y, loss_fn)           # prediction and loss terms (through the projected input)
        L, R, U, lam, g = prob["keep"]
        _, _, S, dev = self._data
        G = self._canon(k)
        rl, ck, lin, rr = G.shape
        owner = self._owner()
        C = self._num_outputs()
        xraw = self._raw_facs[k]
        p_in = xraw.m
        one = ops.ones_factor(G)
        big = 1 << 30
        if C == 1 or owner == k:
            core = G.permute(0, 3, 1, 2).reshape(rl, rr, ck * lin)
            Lf = None if L is None else L.reshape(S, rl)
            if R is None:
                Q = ops.env_update(Lf, one, core.reshape(rl, 1, ck * lin), S, cdiv=big)
            else:
                Q = ops.env_update(Lf, Factor(R.reshape(S, rr), m=rr), core, S)
            Q = Q.view(S, C, lin)
        elif owner < k:
            core = G[:, 0].permute(0, 2, 1)                  # (rl, rr, lin)
            if R is None:
                Q = ops.env_update(L.reshape(S * C, rl), one, core.reshape(rl, 1, lin), S * C, cdiv=big)
            else:
                Q = ops.env_update(L.reshape(S * C, rl), Factor(R.reshape(S, rr), m=rr), core, S * C, cdiv=C)
            Q = Q.view(S, C, lin)
        else:
            core = G[:, 0].permute(2, 0, 1)                  # (rr, rl, lin)
            if L is None:
                Q = ops.env_update(R.reshape(S * C, rr), one, core.reshape(rr, 1, lin), S * C, cdiv=big)
            else:
                Q = ops.env_update(R.reshape(S * C, rr), Factor(L.reshape(S, rl), m=rl), core, S * C, cdiv=C)
            Q = Q.view(S, C, lin)
        V = lam.shape[1]
        if C == 1:
            w = (lam.reshape(S, V) * U.reshape(S, V) ** 2).sum(dim=1).contiguous()
            fq = Factor(Q.view(S, lin), m=lin)
            out = dict(gram=(fq, xraw, one), gw=w, grows=S, rhs=(fq, xraw, one), rw=g.reshape(S).contiguous(), rrows=S)
        else:
            F, Gr = ops.class_rows(Q.contiguous(), U.contiguous(), g)
            xv = Factor(xraw.tensor, m=xraw.m, div=V, map_kind=xraw.map_kind, col=xraw.col)
            out = dict(gram=(Factor(F, m=lin), xv, one), gw=lam.reshape(S * V).contiguous(), grows=S * V,
                       rhs=(Factor(Gr, m=lin), xraw, one), rw=None, rrows=S)
        out.update(m_pos=(lin, p_in, 1), yhat=prob["yhat"], loss=prob["loss"], keep=(prob["keep"], Q))
        return out

    def _set_linear(self, k, new_tensor):
        st = self._plan()[k]
        st.linear.tensor = new_tensor
        self._rebind(self._data[0])          # re-project the bound data; drops every cached environment
        self._stamps = self._stamp()

    def _one_linear_update(self, k, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss):
        self._check_external()
        prob = self._linear_problem(k, y, loss_fn)
        M, b, role_of_pos = self._accumulate(prob, self._gram_mode_for(method, eps, prob))
        W = self._plan()[k].linear.tensor
        try:
            step = self._solve_flat(W.contiguous().view(-1), M, b, prob["m_pos"], role_of_pos, method, eps, prob=prob)
        except _NeedExactGram:
            self.solve_stats["gram_fp64_fallback"] += 1
            del M, b
            M, b, role_of_pos = self._accumulate(prob, "fp64")
            step = self._solve_flat(W.contiguous().view(-1), M, b, prob["m_pos"], role_of_pos, method, eps, prob=prob)
        new = W.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.contiguous().view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
        self._set_linear(k, new)
        if not need_loss:
            return None
        S = prob["yhat"].shape[0]
        return batch_mean_of_means(prob["loss"], batch_size, row_offset=self.shard_offset,
                                   n_total=self.shard_total if self.process_group is not None else S,
                                   group=self.process_group)

    def _gradient_update(self, k, y, loss_fn, lr, batch_size, adaptive_step, max_norm, need_loss):
        """``method='gradient'`` (reference network.py:458-470): minibatch by minibatch, theta += lr * J^T g with the prediction,
        g and J of that minibatch at the CURRENT theta -- the sign is the reference's (ascent for lr > 0).  The environments do
        not depend on the core being updated, so they are built once; per minibatch this is one prediction and one
        right-hand-side pass over its rows.  Under sharding the minibatches are ranges of global rows and each rank adds the
        part it owns."""
        self._check_external()
        node = self.main_nodes[k]
        S = self._data[2]
        dev = self._data[3]
        group = self.process_group
        N = self.shard_total if group is not None else S
        bs = N if batch_size <= 0 else batch_size
        nb = (N + bs - 1) // bs
        P = node.tensor.numel()
        total = 0.0
        for bi in range(nb):
            lo = max(bi * bs, self.shard_offset) - self.shard_offset
            hi = min(min((bi + 1) * bs, N), self.shard_offset + S) - self.shard_offset
            buf = torch.zeros((P + 2,), dtype=torch.float64, device=dev)       # [b | sum of row losses | rows]
            if hi > lo:
                prob = self._site_problem(k, y, loss_fn, rows=(lo, hi))
                rf = prob["rhs"]
                ops.rhs(rf[0], rf[1], rf[2], prob["rw"], prob["rrows"], b=buf[:P])
                rl = prob["loss"].reshape(hi - lo, -1).mean(dim=1)
                buf[P] = rl.sum()
                buf[P + 1] = float(hi - lo)
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(buf, group=group)
            step = self._from_canon(k, buf[:P].reshape(self._canon(k).shape))
            new = node.tensor.detach().clone().contiguous()
            ops.update_node(new.view(-1), step.contiguous().view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
            node.tensor = new
            self._core_changed(k)
            if need_loss:
                total += float((buf[P] / buf[P + 1]).item())
        return torch.tensor(total / nb, dtype=torch.float64) if need_loss else None

    def _gradient_full_update(self, k, y, loss_fn, lr, batch_size, adaptive_step, max_norm, need_loss):
        """``method='gradient'`` in the second (returning) half of a sweep: the reference has no per-minibatch branch there
        (network.py:558-584), so b is accumulated over all minibatches at fixed theta and the step is the ``-b`` of
        ``solve_system`` (:321-322) -- one full-batch step, of the opposite sign to the first half's.  A is not needed."""
        self._check_external()
        prob = self._site_problem(k, y, loss_fn)
        rf = prob["rhs"]
        b = ops.rhs(rf[0], rf[1], rf[2], prob["rw"], prob["rrows"])
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(b, group=self.process_group)
        step = self._from_canon(k, (-b).reshape(self._canon(k).shape))
        node = self.main_nodes[k]
        new = node.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.contiguous().view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
        node.tensor = new
        self._core_changed(k)
        if not need_loss:
            return None
        S = prob["yhat"].shape[0]
        return batch_mean_of_means(prob["loss"], batch_size, row_offset=self.shard_offset,
                                   n_total=self.shard_total if self.process_group is not None else S, group=self.process_group)

    def _update_node(self, node, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss):
        kl = self._linear_site(node)
        if method.lower().startswith("gradient"):
            if kl is not None:
                raise NotImplementedError("method='gradient' on a linear-projection node")
            k = self.main_nodes.index(node)
            if method.lower() == "gradient@batch":
                return self._gradient_update(k, y, loss_fn, lr, batch_size, adaptive_step, max_norm, need_loss)
            return self._gradient_full_update(k, y, loss_fn, lr, batch_size, adaptive_step, max_norm, need_loss)
        if kl is not None:
            return self._one_linear_update(kl, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss)
        return self._one_update(self.main_nodes.index(node), y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm,
                                need_loss)

    def accumulating_swipe(self, x, y_true, loss_fn, node_order=None, batch_size=-1, num_swipes=1, lr=1.0, method="exact",
                           eps=1e-12, eps_decay=None, convergence_criterion=None, orthonormalize=False, verbose=False,
                           skip_second=False, blocks_input=False, timeout=None, data_device=None, model_device=None,
                           disable_tqdm=None, block_callback=None, loss_callback=None, direction="l2r",
                           update_or_reset_stack="reset", adaptive_step=False, min_norm=None, max_norm=None,
                           eps_per_node=False):
        """Reference tensor/network.py:379-608, same keywords and return value.

        Differences that do not change results: A and b of a node are built from the whole data set in
        one pass (they are sums over minibatches at fixed cores); ``batch_size`` only shapes the
        reported mean-of-batch-means loss.  ``update_or_reset_stack`` is accepted; environments are
        always kept incrementally.  ``method='gradient'`` walks the minibatches one by one in the first half of a sweep
        (theta += lr * J^T g per minibatch, network.py:458-470) and takes one full-batch step theta -= lr * b in the second
        (:558-584, :321-322), see ``_gradient_update`` / ``_gradient_full_update``.
        """
        if blocks_input:
            raise NotImplementedError("blocks_input (compressed-data experiment) is outside the sweep path")
        if method == "gradient" and not self._supports_gradient:
            raise NotImplementedError(f"method='gradient' is provided for tensor-train cores only, not for {type(self).__name__}")
        x, y = self._prepare_data(x, y_true, data_device, model_device)
        # node lists of the two half-sweeps, exactly as network.py:418-425,520-527 derive them
        if node_order is None:
            first, second = list(self.train_nodes), list(self.train_nodes)
        elif isinstance(node_order, tuple):
            first, second = list(node_order[0]), list(node_order[1])
        else:
            first, second = list(node_order), list(reversed(list(node_order)))
        first = first if direction == "l2r" else list(reversed(first))
        second = second if direction == "r2l" else list(reversed(second))
        halves = (first, second)
        col = lambda nd, i: self.node_indices[nd] if nd in self.node_indices else ("x", id(nd))
        sched = sweep_schedule([col(nd, i) for i, nd in enumerate(first)], [col(nd, i) for i, nd in enumerate(second)],
                               num_swipes, eps, eps_decay, skip_second, direction, eps_per_node)
        start = time.time() if timeout is not None else None
        need_loss = loss_callback is not None or (verbose and verbose > 1)
        for NS, half, pos, eps_ in sched:
            node = halves[half][pos]
            if isinstance(eps_, Exception):
                raise eps_                  # an epsilon list too short for this entry: the reference's IndexError, at the same point
            if timeout is not None and (time.time() - start) > timeout:
                print(f"Timeout reached ({timeout} seconds). Stopping accumulating_swipe.")
                return False
            _method = "exact" if (eps_ == 0 and method == "ridge_exact") else method
            if method == "gradient" and half == 0:
                _method = "gradient@batch"      # per-minibatch steps exist in the first half only (network.py:469-470 vs :558-584)
            try:
                loss = self._update_node(node, y, loss_fn, _method, eps_, lr, batch_size, adaptive_step, max_norm, need_loss)
            except torch.linalg.LinAlgError:
                if verbose and verbose > 0:
                    print(f"Singular system for node {node.name}")
                return False
            going_right = half == 0     # the first half re-gauges to the left, the second to the right (network.py:487-488,585-586)
            if orthonormalize and node not in self.main_nodes:
                raise NotImplementedError("QR re-gauge of a linear-projection node")
            if orthonormalize:
                if going_right:
                    self.node_orthonormalize_left(node)
                else:
                    self.node_orthonormalize_right(node)
            if self.lean_envs and node in self.main_nodes:
                k = self.main_nodes.index(node)
                drop = self._right if going_right else self._left
                for j in [j for j in drop if (j <= k + 1 if going_right else j >= k - 1)]:
                    del drop[j]
            if need_loss:
                lv = float(loss.item())
                if verbose and verbose > 1:
                    print(f"NS: {NS}, {'Left' if half == 0 else 'Right'} loss ({node.name}):", lv, f" (eps: {eps_})")
                if loss_callback is not None:
                    loss_callback(NS, node, lv)
            if convergence_criterion is not None and convergence_criterion():
                if verbose and verbose > 0:
                    print("Converged (left pass)" if half == 0 else "Converged (right pass)")
                if block_callback is not None:
                    block_callback(NS, node)
                return True
            if block_callback is not None:
                block_callback(NS, node)
        return True

    # ------------------------------------------------------------------ matrix-free sweeps
    def _krylov_setup(self, k, y, loss_fn, prob=None):
        prob = self._site_problem(k, y, loss_fn) if prob is None else prob
        gf, rf = prob["gram"], prob["rhs"]
        b = ops.rhs(rf[0], rf[1], rf[2], prob["rw"], prob["rrows"])
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(b, group=self.process_group)
        # v -> J^T H J v on the virtual rows: the built-in operator of the on-device Krylov drivers (csrc/krylov.cu); calling the
        # object applies it to a tensor (SciPy bridge)
        m_pos = prob["m_pos"]
        op = ops.Operator(m_pos[0] * m_pos[1] * m_pos[2], factors=gf, w=prob["gw"], rows=prob["grows"], group=self.process_group)
        op.keep = prob["keep"]
        return prob, b, op

    def _krylov_problem(self, node, y, loss_fn):
        """(per-row loss, right-hand side b, matvec v -> J^T H J v) of one node, everything flat in canonical order."""
        kl = self._linear_site(node)
        if kl is not None:
            prob, b, matvec = self._krylov_setup(kl, y, loss_fn, prob=self._linear_problem(kl, y, loss_fn))
            return prob["loss"], b, matvec
        prob, b, matvec = self._krylov_setup(self.main_nodes.index(node), y, loss_fn)
        return prob["loss"], b, matvec

    def _apply_step(self, node, step_c, lr):
        kl = self._linear_site(node)
        if kl is not None:
            new = node.tensor.detach().clone().contiguous()
            ops.update_node(new.view(-1), step_c.contiguous().view(-1), lr=lr)
            self._set_linear(kl, new)
            return
        k = self.main_nodes.index(node)
        step = self._from_canon(k, step_c.reshape(self._canon(k).shape))
        new = node.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.view(-1), lr=lr)
        node.tensor = new
        self._core_changed(k)

    def _krylov_swipe(self, x, y_true, loss_fn, solve, batch_size, num_swipes, lr, timeout, data_device, model_device,
                      block_callback, loss_callback, what):
        x, y = self._prepare_data(x, y_true, data_device, model_device)
        start = time.time() if timeout is not None else None
        for NS in range(num_swipes):
            order = list(self.train_nodes) if NS % 2 == 0 else list(reversed(self.train_nodes))
            for node in order:
                if timeout is not None and (time.time() - start) > timeout:
                    print(f"Timeout reached ({timeout} seconds). Stopping {what}.")
                    return False
                self._check_external()
                loss_rows, b, matvec = self._krylov_problem(node, y, loss_fn)
                if loss_callback is not None:
                    S = loss_rows.shape[0]
                    lv = batch_mean_of_means(loss_rows, batch_size, row_offset=self.shard_offset,
                                             n_total=self.shard_total if self.process_group is not None else S,
                                             group=self.process_group)
                    loss_callback(float(lv.item()))
                step_c = solve(node, matvec, b)                       # canonical order, flat
                self._apply_step(node, step_c, lr)
                if block_callback is not None:
                    block_callback(NS, node)
        return True

    def lanczos_swipe(self, x, y_true, loss_fn, batch_size=1, num_swipes=1, lr=1.0, max_iter=50, tol=1e-6, verbose=False,
                      timeout=None, data_device=None, model_device=None, disable_tqdm=None, block_callback=None,
                      loss_callback=None, x0_fn=None):
        """Reference tensor/network.py:709-832: Lanczos-Galerkin solve of A step = -b per node, matrix-free.
        The start vector is random there (``randn_like``, :793); ``x0_fn(node, b)`` lets a caller inject one."""

        def solve(node, matvec, b):
            from ..krylov import lanczos
            rhs = -b
            x0 = x0_fn(node, b) if x0_fn is not None else torch.randn_like(rhs)
            # r0 = rhs - A x0, max_iter Lanczos vectors (stop when |w_j| < tol), x = x0 + V T^-1 |r0| e1: the recurrence of
            # network.py:793-824 inside libtn_b200.so (tn_lanczos), scalars on the device
            return lanczos(matvec, rhs.contiguous().view(-1), x0.contiguous().reshape(-1), maxiter=max_iter, tol=tol)

        return self._krylov_swipe(x, y_true, loss_fn, solve, batch_size, num_swipes, lr, timeout, data_device, model_device,
                                  block_callback, loss_callback, "lanczos_swipe")

    def scipy_swipe(self, x, y_true, loss_fn, solver, batch_size=1, num_swipes=1, lr=1.0, max_iter=50, tol=1e-6, verbose=False,
                    timeout=None, data_device=None, model_device=None, disable_tqdm=None, block_callback=None,
                    loss_callback=None):
        """Reference tensor/network.py:834-932.  ``solver`` is ``scipy.sparse.linalg.cg``/``minres``-like or one of
        the strings 'cg' / 'minres', which select the on-device float64 solvers of this package
        (the reference runs the SciPy recurrences in float32 on the host, :918,921)."""
        sols = getattr(self, "_node_sols", None)
        if sols is None:
            sols = self._node_sols = {}

        def solve(node, matvec, b):
            from ..krylov import cg, minres, scipy_bridge
            prev = sols.get(node)
            if isinstance(solver, str):
                fn = {"cg": cg, "minres": minres}[solver]
                xs = fn(matvec, -b, x0=prev, maxiter=max_iter, rtol=tol)
            else:
                xs = scipy_bridge(solver, matvec, -b, x0=prev, maxiter=max_iter, rtol=tol)
            sols[node] = xs
            return xs if isinstance(xs, torch.Tensor) else torch.as_tensor(xs, dtype=b.dtype, device=b.device)

        return self._krylov_swipe(x, y_true, loss_fn, solve, batch_size, num_swipes, lr, timeout, data_device, model_device,
                                  block_callback, loss_callback, "scipy_swipe")


class _NeedExactGram(Exception):
    """The refinement of a tensor-core Gram mode did not reach its residual: redo the site with the fp64 Gram."""


class _FixedTerms:
    """Adapter: a (grad, hessian) pair supplied by the caller, presented as a loss object."""

    def __init__(self, g, H):
        self.g, self.H = g, H

    def forward(self, y_pred, y):
        return torch.zeros(self.g.shape[0], dtype=self.g.dtype, device=self.g.device), self.g, self.H


class SumOfNetworks(TensorNetwork):
    """Sum of several networks fed by (feature slices of) the same input -- the reference's "type-I" models
    (tensor/network.py:988-1060; built by models/tensor_train.py:138-189 as chains of 1..N cores).

    The prediction is the sum of the members' predictions.  A node is updated inside its own member with the other
    members' outputs held fixed (added to the member's prediction before the loss is evaluated), which is exactly what
    the reference's inherited sweep does through ``forward`` (sum) + ``get_A_b`` (delegated to the member).
    """

    def __init__(self, networks, output_labels=("s",), sample_dim="s", train_operators=True):
        input_nodes, main_nodes, train_nodes = [], [], []
        for i, net in enumerate(networks, 1):
            for n in net.input_nodes:
                n.name = f"{n.name}_n{i}"
            for n in net.main_nodes:
                n.name = f"{n.name}_n{i}"
            input_nodes.extend(net.input_nodes)
            main_nodes.extend(net.main_nodes)
            train_nodes.extend(net.train_nodes if train_operators else net.main_nodes)
        super().__init__(input_nodes, main_nodes, train_nodes, output_labels=output_labels, sample_dim=sample_dim)
        self.networks = list(networks)
        self._member_pred = {}

    def _plan(self):
        raise NotImplementedError("SumOfNetworks delegates to its members")

    def _member_of(self, node):
        for j, net in enumerate(self.networks):
            if any(node is n for n in net.main_nodes) or any(node is n for n in net.train_nodes):
                return j, net
        raise ValueError("Node not found in any network")

    def _member_width(self, net):
        s = net._plan()[0] if not hasattr(net, "_rank") else None
        if s is not None:
            if getattr(s, "linear", None) is not None:
                return s.linear.tensor.shape[1]          # members built from linear-projection trains see the raw features
            return net._phys_size(s)
        return net._plan()[0].dim_size("p")

    def _member_input(self, net, x):
        if hasattr(net, "_columns"):
            # conv-TT member (AAMNST.py:160-168): members after the first see x[:, :patches-1, :pixels-1], i.e. the leading part
            # of every non-sample axis up to the size its input nodes were built with (reference network.py:1012)
            want = tuple(net.input_nodes[0].tensor.shape[1:])
            if x.dim() != 1 + len(want) or any(w > h for w, h in zip(want, x.shape[1:])):
                raise ValueError(f"input of shape {tuple(x.shape)} cannot feed a member built for {want} per sample")
            return x if tuple(x.shape[1:]) == want else x[(slice(None),) + tuple(slice(0, w) for w in want)]
        f = self._member_width(net)
        return x if x.shape[1] == f else x[:, :f]

    def forward(self, x, to_tensor=False):
        y = None
        for net in self.networks:
            yj = net._chain_forward(self._member_input(net, x))
            y = yj if y is None else y + yj
        # the members' output legs survive even when the sum itself was declared with ('s',) only (AAMNST.py:168 does that;
        # the reference's forward only moves the listed labels to the front, network.py:1013-1015)
        first = self.networks[0]
        out_labels = [l for l in self.output_labels if l != self.sample_dim] or [l for l in first.output_labels if l != first.sample_dim]
        if not out_labels:
            y = y[:, 0]
        return y if to_tensor else TensorNode(y, [self.sample_dim] + out_labels, name="O")

    def reset_stacks(self, node=None):
        for net in self.networks:
            net.reset_stacks()
        self._member_pred = {}

    def set_input(self, x):
        key = self._key_of(x)
        if key == self._data_key:
            return False
        self._data_key = key
        self._data = (x, [self._member_input(net, x) for net in self.networks])
        for net, xj in zip(self.networks, self._data[1]):
            net.set_input(xj)
        self._member_pred = {}
        return True

    def _require_cuda(self, dev):
        for net in self.networks:
            net._require_cuda(dev)

    def _update_node(self, node, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss):
        j, net = self._member_of(node)
        xs = self._data[1]
        offset = None
        for i, other in enumerate(self.networks):
            if i == j:
                continue
            stamp = other._stamp() if hasattr(other, "_stamp") else None
            cached = self._member_pred.get(i)
            if cached is None or cached[0] != stamp:
                cached = (stamp, other._chain_forward(xs[i]))
                self._member_pred[i] = cached
            offset = cached[1] if offset is None else offset + cached[1]
        net.process_group, net.shard_offset, net.shard_total = self.process_group, self.shard_offset, self.shard_total
        net.gram_mode = self.gram_mode
        net._yhat_offset = offset
        if method.lower().startswith("gradient") and not net._supports_gradient:
            raise NotImplementedError(f"method='gradient' is provided for tensor-train cores only, not for {type(net).__name__}")
        try:
            return net._update_node(node, y, loss_fn, method, eps, lr, batch_size, adaptive_step, max_norm, need_loss)
        finally:
            net._yhat_offset = None

    def orthonormalize_left(self):
        for net in self.networks:
            net.orthonormalize_left()

    def orthonormalize_right(self):
        for net in self.networks:
            net.orthonormalize_right()

    def node_orthonormalize_left(self, node):
        self._member_of(node)[1].node_orthonormalize_left(node)

    def node_orthonormalize_right(self, node):
        self._member_of(node)[1].node_orthonormalize_right(node)

    def get_A_b(self, node, grad=None, hessian=None, method=None, y=None, loss_fn=None):
        return self._member_of(node)[1].get_A_b(node, grad, hessian, method=method, y=y, loss_fn=loss_fn)

    def lanczos_swipe(self, *a, **k):
        raise NotImplementedError("matrix-free sweeps over a SumOfNetworks")

    scipy_swipe = lanczos_swipe
