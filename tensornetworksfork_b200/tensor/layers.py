"""Layer constructors: topology and initial state of the models the sweep fits.

They mirror the reference's constructors (tensor/layers.py:9-97 MainNodeLayer, :114-192
TensorNetworkLayer, :194-221 TensorTrainLayer, :1549-1625 CPDLayer) in name, arguments, label
scheme and -- because parity needs identical initial cores -- in the *order and shape of the random
draws* for a given seed.  The graphs they build are handed to the B200 sweep engine
(``network.TensorNetwork`` / ``cpd.CPDNetwork``) instead of the reference's einsum engine.
"""
import torch
from torch import nn

from .network import TensorNetwork
from .node import TensorNode


def chain_bond_dims(n_sites, r, f, constrict_bond=True, perturb=False):
    """Bond dimensions d[0..n] of a chain (d[0] = d[n] = 1).

    Random-init chains grow bonds from both ends towards the middle, ``min(r, d*f)`` per step when
    ``constrict_bond`` (reference layers.py:20-30,58-73: left side gets the extra core for odd
    remainders).  Identity-init ('perturb') chains grow from the left only and keep the last bond at
    ``r`` (layers.py:40-57).
    """
    if n_sites == 1:
        return [1, 1]
    grow = (lambda d: min(r, d * f)) if constrict_bond else (lambda d: r)
    d = [None] * (n_sites + 1)
    d[0] = d[n_sites] = 1
    if perturb:
        d[n_sites - 1] = r
        for j in range(n_sites - 2):
            d[j + 1] = grow(d[j])
        if n_sites == 2:
            d[1] = grow(1)
        return d
    n_left = 1 + max(0, (n_sites - 2) // 2)
    n_right = 1 + max(0, (n_sites - 3) // 2)
    for j in range(n_left):
        d[j + 1] = grow(d[j])
    for j in range(n_right):
        d[n_sites - 1 - j] = grow(d[n_sites - j])
    return d


def _identity_core(rl, f, rr, dtype=None):
    """Core that passes the bond through on the LAST feature (the bias column) and is zero elsewhere
    (reference layers.py:32-38): ones if either bond is 1, else a rectangular identity."""
    blk = torch.ones(rl, rr, dtype=dtype) if (rl == 1 or rr == 1) else torch.eye(rl, rr, dtype=dtype)
    core = torch.zeros(rl, 1, f, rr, dtype=blk.dtype)
    core[:, 0, f - 1, :] = blk
    return core


class MainNodeLayer(nn.Module):
    """The train cores A1..AN with labels r{i} -[c{i}|c, p{i}]- r{i+1} (reference layers.py:9-97)."""

    def __init__(self, N, r, f, output_shape=tuple(), down_label="p", horizontal_label="r{0}", constrict_bond=True,
                 perturb=False, dtype=None):
        super().__init__()
        output_shape = output_shape if isinstance(output_shape, tuple) else (output_shape,)
        if N == 1:
            r = 1
        dims = chain_bond_dims(N, r, f, constrict_bond, perturb)
        blocks = None
        if perturb:
            # draw order matters for seed parity: first core's noise, then last core's (layers.py:41-44)
            first = _identity_core(1, f, dims[1], dtype)
            first = first * (1 + 0.02 * torch.randn_like(first))
            last = _identity_core(dims[N - 1] if N > 1 else r, f, 1, dtype)
            last = last * (1 + 0.02 * torch.randn_like(last))
            blocks = [first] + [_identity_core(dims[i], f, dims[i + 1], dtype) for i in range(1, N - 1)] + ([last] if N > 1 else [])
        self.labels = ["s"]
        self.nodes = []
        for i in range(1, N + 1):
            if i - 1 < len(output_shape):
                up, up_label = output_shape[i - 1], f"c{i}"
                self.labels.append(up_label)
            else:
                up, up_label = 1, "c"
            left_label, right_label = horizontal_label.format(i), horizontal_label.format(i + 1)
            init = blocks[i - 1] if perturb else (dims[i - 1], up, f, dims[i])
            self.nodes.append(TensorNode(init, [left_label, up_label, down_label.format(i), right_label], l=left_label,
                                         r=right_label, name=f"A{i}", dtype=dtype))


class InputNodeLayer(nn.Module):
    """Placeholder input nodes X1..XN with labels (s, p{i}) (reference layers.py:99-112)."""

    def __init__(self, N, f, label="p", dtype=None):
        super().__init__()
        self.nodes = [TensorNode((1, f), ["s", label.format(i)], name=f"X{i}", dtype=dtype) for i in range(1, N + 1)]


class TensorNetworkLayer(nn.Module):
    """Owner of a tensor network; state = the tensors of its train nodes (reference layers.py:114-192)."""

    def __init__(self, tensor_network=None):
        super().__init__()
        self.set_tensor_network(tensor_network)

    def set_tensor_network(self, tensor_network=None):
        self.tensor_network = tensor_network
        self.labels = tensor_network.output_labels if tensor_network is not None else None
        self.parametrized = False
        self.nodes = tensor_network.train_nodes if tensor_network is not None else []

    def node_states(self, detach=True):
        return {f"tensor_param_{i}": (n.tensor.detach().clone() if detach else n.tensor)
                for i, n in enumerate(self.tensor_network.train_nodes)}

    def load_node_states(self, tensor_params, set_value=False):
        for i, n in enumerate(self.tensor_network.train_nodes):
            key = f"tensor_param_{i}"
            if key not in tensor_params:
                raise ValueError(f"Missing parameter: {key}")
            if set_value:
                n.tensor = tensor_params[key]
            else:
                n.tensor.data.copy_(tensor_params[key].detach().clone())
        self.tensor_network.reset_stacks()

    def cuda(self, *a, **k):
        self.tensor_network.cuda()
        return super().cuda(*a, **k)

    def to(self, *a, **k):
        self.tensor_network.to(*a, **k)
        return super().to(*a, **k)

    def cpu(self, *a, **k):
        self.tensor_network.to("cpu")
        return super().cpu(*a, **k)

    def forward(self, x, to_tensor=True):
        out = self.tensor_network.forward(x)
        if self.labels is not None:
            out.permute_first(*self.labels)
        return out.tensor if to_tensor else out

    def num_parameters(self):
        return sum(n.tensor.numel() for n in self.tensor_network.train_nodes)

    @staticmethod
    def zip_connect(nodes1, nodes2, label="p", priority=-1):
        if len(nodes1) != len(nodes2):
            raise ValueError("The number of nodes in both lists must be the same.")
        for i, (a, b) in enumerate(zip(nodes1, nodes2), 1):
            a.connect(b, label.format(i), priority=priority)

    @staticmethod
    def horizontal_connect(nodes):
        for a, b in zip(nodes[:-1], nodes[1:]):
            if a.right_labels and b.left_labels and a.right_labels[0] != b.left_labels[0]:
                raise ValueError(f"Right label of the first node does not match left label of the second node. "
                                 f"Nodes: {a.name}, {b.name}")
            a.connect(b, a.right_labels[0], priority=1)


class TensorTrainLayer(TensorNetworkLayer):
    """Tensor train with one input per core (reference layers.py:194-221)."""

    def __init__(self, num_carriages, bond_dim, input_features, output_shape=tuple(), squeeze=True, constrict_bond=True,
                 perturb=False, dtype=None, seed=None):
        super().__init__()
        self.num_carriages = num_carriages
        self.bond_dim = bond_dim
        self.input_features = input_features
        self.output_shape = output_shape if isinstance(output_shape, tuple) else (output_shape,)
        if seed is not None:
            torch.manual_seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed(seed)
        self.main_node_layer = MainNodeLayer(num_carriages, bond_dim, input_features, output_shape=output_shape,
                                             down_label="p{0}", constrict_bond=constrict_bond, perturb=perturb, dtype=dtype)
        self.horizontal_connect(self.main_node_layer.nodes)
        self.input_node_layer = InputNodeLayer(num_carriages, input_features, label="p{0}", dtype=dtype)
        self.zip_connect(self.input_node_layer.nodes, self.main_node_layer.nodes, label="p{0}")
        if squeeze:
            for n in self.main_node_layer.nodes:
                n.squeeze(self.main_node_layer.labels)
        self.set_tensor_network(TensorNetwork(self.input_node_layer.nodes, self.main_node_layer.nodes,
                                              output_labels=self.main_node_layer.labels))


class TensorTrainLinearLayer(TensorNetworkLayer):
    """Tensor train whose cores see a trainable linear projection of the input: per column a core A_k over `linear_dim` and a
    projection L_k (linear_dim x input_features); train-node order A1, L1, A2, L2, ... (reference layers.py:308-343; same
    constructor, labels, names and random draws)."""

    def __init__(self, num_carriages, bond_dim, input_features, linear_dim, output_shape=tuple(), squeeze=True, constrict_bond=True,
                 perturb=False, dtype=None, seed=None):
        super().__init__()
        self.num_carriages = num_carriages
        self.bond_dim = bond_dim
        self.input_features = input_features
        self.output_shape = output_shape if isinstance(output_shape, tuple) else (output_shape,)
        self.linear_dim = linear_dim
        if seed is not None:
            torch.manual_seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed(seed)
        self.main_node_layer = MainNodeLayer(num_carriages, bond_dim, linear_dim, output_shape=output_shape, down_label="lin{0}",
                                             constrict_bond=constrict_bond, perturb=perturb, dtype=dtype)
        self.horizontal_connect(self.main_node_layer.nodes)
        lin_nodes = [TensorNode((linear_dim, input_features), [f"lin{i}", f"p{i}"], name=f"L{i}", dtype=dtype)
                     for i in range(1, num_carriages + 1)]
        self.linear_layer = nn.Module()
        self.linear_layer.nodes = lin_nodes
        self.zip_connect(self.main_node_layer.nodes, lin_nodes, label="lin{0}", priority=2)
        self.input_node_layer = InputNodeLayer(num_carriages, input_features, label="p{0}", dtype=dtype)
        self.zip_connect(lin_nodes, self.input_node_layer.nodes, label="p{0}", priority=1)
        if squeeze:
            for n in self.main_node_layer.nodes:
                n.squeeze(self.main_node_layer.labels)
        train = [n for col in zip(self.main_node_layer.nodes, lin_nodes) for n in col]
        self.set_tensor_network(TensorNetwork(self.input_node_layer.nodes, main_nodes=self.main_node_layer.nodes, train_nodes=train,
                                              output_labels=self.main_node_layer.labels))


def get_cum_sum_operator(n, num_carriages, input_features, dtype=None):
    """Operator node of site ``n`` of the reference's cum-sum train as a dense tensor (left, p_in, p_core, right), reference
    tensor/layers.py:408-423: entry (i, k, k, m) is one where i <= k (a single all-ones row on the first site) and m = k (m = 0 on the
    last site).  ``CumSumLayer`` here never builds it -- ``tensor/cumsum.py`` uses the closed form of the contracted train -- it is kept
    for callers that assemble the operator train themselves."""
    f = input_features
    left = 1 if n == 0 else f
    right = 1 if n == num_carriages - 1 else f
    op = torch.zeros((left, f, f, right), dtype=dtype)
    for k in range(f):
        op[: (1 if left == 1 else k + 1), k, k, 0 if right == 1 else k] = 1
    return op


class CumSumLayer(TensorNetworkLayer):
    """Tensor train over ordered feature tuples (reference layers.py:425-477).  Same cores and draws as
    ``TensorTrainLayer``; the reference's dense cum-sum operator nodes are replaced by the closed form in
    ``cumsum.CumSumNetwork``.  Like the reference, ``seed`` is accepted and not applied here."""

    def __init__(self, num_carriages, bond_dim, input_features, output_shape=tuple(), squeeze=True, constrict_bond=True,
                 perturb=False, dtype=None, seed=None):
        from .cumsum import CumSumNetwork
        super().__init__()
        self.num_carriages = num_carriages
        self.input_features = input_features
        self.main_node_layer = MainNodeLayer(num_carriages, bond_dim, input_features, output_shape=output_shape,
                                             down_label="p{0}", constrict_bond=constrict_bond, perturb=perturb, dtype=dtype)
        self.horizontal_connect(self.main_node_layer.nodes)
        self.input_node_layer = InputNodeLayer(num_carriages, input_features, label="p{0}", dtype=dtype)
        self.zip_connect(self.input_node_layer.nodes, self.main_node_layer.nodes, label="p{0}", priority=1)
        if squeeze:
            for n in self.main_node_layer.nodes:
                n.squeeze(self.main_node_layer.labels)
        self.set_tensor_network(CumSumNetwork(self.input_node_layer.nodes, self.main_node_layer.nodes,
                                              output_labels=self.main_node_layer.labels))


class CPDLayer(TensorNetworkLayer):
    """Rank-R canonical polyadic model: factor i is (b, p[, o]) (reference layers.py:1549-1625)."""

    def __init__(self, num_factors, rank, input_features, output_shape=tuple(), perturb=False, seed=None):
        from .cpd import CPDNetwork
        self.num_factors = num_factors
        self.rank = rank
        self.input_features = input_features
        self.output_shape = output_shape if isinstance(output_shape, tuple) else (output_shape,)
        if seed is not None:
            torch.manual_seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed(seed)
        # inputs are created first: their (unused) random draws advance the generator (layers.py:1569-1576)
        x_nodes = [TensorNode((1, input_features), ["s", "p"], name=f"X{i}") for i in range(1, num_factors + 1)]
        factors = []
        labels = ["s"]
        for i in range(1, num_factors + 1):
            out_dim = self.output_shape[i - 1] if i - 1 < len(self.output_shape) else 1
            if i == 1:
                if num_factors == 1:
                    node = TensorNode((input_features, out_dim), ["p", "o"], name="A1")
                else:
                    node = TensorNode((rank, input_features, out_dim), ["b", "p", "o"], name="A1")
                labels.append("o")
            else:
                init = (rank, input_features)
                if perturb:
                    bias = torch.ones(rank, 1)
                    if i == num_factors:
                        bias = bias + 0.02 * torch.randn(rank, 1)
                    init = torch.cat((torch.zeros(rank, input_features - 1), bias), dim=1)
                node = TensorNode(init, ["b", "p"], name=f"A{i}")
            factors.append(node)
        for xn, an in zip(x_nodes, factors):
            xn.connect(an, "p")
        self.x_nodes = x_nodes
        super().__init__(CPDNetwork(x_nodes, factors, output_labels=tuple(labels), sample_dim="s"))
        self.nodes = factors
        self.labels = tuple(labels)


class TensorTrainDMRGInfiLayer(TensorNetworkLayer):
    """Two-site growth of a tensor train ("infinite DMRG" style; reference layers.py:480-680).

    Starts as a 2-core chain AL1-AR1.  ``grow_middle`` inserts a trainable 2-site block
    ``D[r_l, pL, pR, r_r]`` fed by two inputs between the two middle cores (only the block is trained by
    the next sweep); ``split_node`` splits it back into two cores by a truncated SVD on the host.
    """

    def __init__(self, bond_dim, input_features, output_shape=tuple(), ring=False, squeeze=True, constrict_bond=True):
        self.num_carriages = 2
        self.bond_dim = bond_dim
        self.input_features = input_features
        self.output_shape = output_shape if isinstance(output_shape, tuple) else (output_shape,)
        self.ring = ring
        # draw order (inputs first, then the cores) follows the reference for seed parity
        self.x_nodes = [TensorNode((1, input_features), ["s", "pL1"], name="XL1"),
                        TensorNode((1, input_features), ["s", "pR1"], name="XR1")]
        bond = min(bond_dim, input_features) if constrict_bond else bond_dim
        self.ranks = [(1, bond), (bond, 1)]
        self.labels = ["s", "c1"]
        n1 = TensorNode((self.output_shape[0], input_features, bond), ["c1", "pL1", "r1"], r="r1", name="AL1")
        n1.connect(self.x_nodes[0], "pL1", priority=2)
        n2 = TensorNode((bond, input_features), ["r1", "pR1"], l="r1", name="AR1")
        n2.connect(self.x_nodes[1], "pR1", priority=2)
        n1.connect(n2, "r1", priority=0)
        self.nodes = [n1, n2]
        if squeeze:
            for n in self.nodes:
                n.squeeze(self.labels)
        super().__init__(TensorNetwork(self.x_nodes, self.nodes, output_labels=self.labels))
        self.nodes = [n1, n2]

    def _rebuild(self, train_nodes, device):
        self.tensor_network = TensorNetwork(self.x_nodes, self.nodes, train_nodes=train_nodes, output_labels=self.labels)
        self.tensor_network.to(device)

    def grow_middle(self):
        """Insert the 2-site block between the two middle cores (reference layers.py:556-614)."""
        n = self.num_carriages
        pl, pr = f"pL{n}", f"pR{n}"
        x1 = TensorNode((1, self.input_features), ["s", pl], name=f"XL{n}")
        x2 = TensorNode((1, self.input_features), ["s", pr], name=f"XR{n}")
        left, right = self.nodes[n // 2 - 1], self.nodes[n // 2]
        old = left.right_labels[0]
        left.connections.pop(old, None)
        right.connections.pop(right.left_labels[0], None)
        ll, rl = old + "L", right.left_labels[0] + "R"
        left.right_labels = [ll]
        left.dim_labels[-1] = ll
        right.left_labels = [rl]
        right.dim_labels[0] = rl
        b1, b2 = left.dim_size(ll), right.dim_size(rl)
        block = TensorNode((b1, 1, self.input_features, self.input_features, b2), [ll, f"c{n}", pl, pr, rl], l=ll, r=rl,
                           name=f"D{n}")
        x1.connect(block, pl)
        x2.connect(block, pr)
        self.x_nodes.insert(n // 2, x2)
        self.x_nodes.insert(n // 2, x1)
        block.connect(left, ll)
        block.connect(right, rl)
        block.squeeze()
        self.nodes.insert(n // 2, block)
        self.num_carriages += 1
        self._rebuild([block], left.tensor.device)

    def split_node(self, left_labels, right_labels, rank, err=None, is_last=False):
        """Split the block by a truncated SVD into two cores (host-side, reference layers.py:616-680).
        Returns the discarded tail of the singular values."""
        import numpy as np
        n = self.num_carriages
        node = self.nodes[n // 2]
        cur_l, cur_r = node.left_labels[0], node.right_labels[0]
        node.permute_first(*left_labels)
        node.permute_last(*right_labels)
        ldims = [node.dim_size(l) for l in left_labels]
        rdims = [node.dim_size(l) for l in right_labels]
        u, sv, v = torch.linalg.svd(node.tensor.reshape(int(np.prod(ldims)), int(np.prod(rdims))), full_matrices=False)
        if is_last:
            v = sv.diag() @ v
        tail = torch.flip(sv, dims=[0]).cumsum(0)
        if err is not None:
            rank = max(min(rank, int((tail > err).sum())), 1)
        split_err = tail[-rank]
        u = u[:, :rank].reshape(ldims + [rank])
        v = v[:rank].reshape([rank] + rdims)
        bond = f"r{n}"
        a = TensorNode(u, list(left_labels) + [bond], r=bond, l=cur_l, name=f"AL{n}")
        b = TensorNode(v, [bond] + list(right_labels), r=cur_r, l=bond, name=f"AR{n}")
        if cur_l in node.connections:
            node.connections[cur_l].connect(a, cur_l)
        if cur_r in node.connections:
            node.connections[cur_r].connect(b, cur_r)
        a.connect(b, bond)
        xa, xb = self.x_nodes[n // 2], self.x_nodes[n // 2 + 1]
        xa.reset_connections()
        xb.reset_connections()
        xa.connect(a, xa.dim_labels[1])
        xb.connect(b, xb.dim_labels[1])
        pos = self.nodes.index(node)
        self.nodes[pos:pos + 1] = [a, b]
        self.num_carriages += 1
        self._rebuild([], node.tensor.device)
        return split_err


class TensorConvolutionTrainLayer(TensorNetworkLayer):
    """Patch/pixel "conv-TT": per column an input node x[s, patches, patch_pixels], a pixel core C_k over the pixels of a
    patch (bond ``convolution_bond``) and a patch core A_k over the patches (bond ``bond_dim``); the output leg sits on A_1.

    Same constructor, node names, labels, train-node order (C1, A1, C2, A2, ...) and random-number consumption as the
    reference (tensor/layers.py:791-890), so the same ``torch.manual_seed`` gives the same initial cores; the network object
    is the B200 engine ``ConvTrainNetwork`` (matrix-free sweeps)."""

    def __init__(self, num_carriages, bond_dim, num_patches, patch_pixels, output_shape, ring=False, convolution_bond=-1, dtype=None,
                 constrict_bond=True, perturb=False):
        from .conv import ConvTrainNetwork
        if ring:
            raise NotImplementedError("Ring structure is not implemented for TensorConvolutionTrainLayer.")
        self.num_carriages = num_carriages
        self.bond_dim = bond_dim
        self.num_patches = num_patches
        self.patch_pixels = patch_pixels
        self.output_shape = output_shape if isinstance(output_shape, tuple) else (output_shape,)
        self.ring = ring
        self.convolution_bond = convolution_bond
        self.output_labels = ("s",)
        n, Q, T, r, CB = num_carriages, num_patches, patch_pixels, bond_dim, convolution_bond

        if perturb:
            # first core random, every other core the identity on the bias patch (reference :812-836)
            def bias_identity(rl, rr):
                blk = torch.diag_embed(torch.ones(rr, dtype=dtype)) if rl == rr else torch.ones(rl, rr, dtype=dtype)
                return torch.cat((torch.zeros(rl, Q - 1, rr), blk.unsqueeze(1)), dim=1)
            first = torch.randn((1, Q, r), dtype=dtype)
            blocks = [first] + [bias_identity(r, r) for _ in range(n - 2)] + [bias_identity(r, 1)]
            blocks = [b.unsqueeze(1) for b in blocks]
        else:
            blocks = [(r if i != 1 else 1, self.output_shape[i - 1] if i <= len(self.output_shape) else 1, Q, r if i != n else 1)
                      for i in range(1, n + 1)]

        x_nodes, conv_blocks, train_blocks = [], [], []
        for i in range(1, n + 1):
            up = f"c{i}" if i - 1 < len(self.output_shape) else "c"
            x_node = TensorNode((1, Q, T), ["s", "patches", "patch_pixels"], name=f"X{i}")
            if CB > 0:
                conv = TensorNode((CB if i != 1 else 1, T, CB if i != n else 1), [f"CB{i}", "patch_pixels", f"CB{i + 1}"],
                                  l=f"CB{i}", r=f"CB{i + 1}", name=f"C{i}")
            else:
                conv = TensorNode((T,), ["patch_pixels"], name=f"C{i}")
            train = TensorNode(blocks[i - 1], [f"r{i}", up, "patches", f"r{i + 1}"], l=f"r{i}", r=f"r{i + 1}", name=f"A{i}")
            x_nodes.append(x_node)
            conv_blocks.append(conv)
            train_blocks.append(train)
            if i < len(self.output_shape) + 1:
                self.output_labels = self.output_labels + (f"c{i}",)

        self.nodes = []
        for xn, cb, tb in zip(x_nodes, conv_blocks, train_blocks):
            xn.connect(tb, "patches")
            cb.connect(xn, "patch_pixels")
            self.nodes.append(cb)
            self.nodes.append(tb)
        for i in range(1, n):
            train_blocks[i - 1].connect(train_blocks[i], f"r{i + 1}")
        if CB > 0:
            for i in range(1, n):
                conv_blocks[i - 1].connect(conv_blocks[i], f"CB{i + 1}")
        for nd in train_blocks + conv_blocks:
            nd.squeeze()

        self.x_nodes = x_nodes
        self.conv_blocks = conv_blocks
        self.train_blocks = train_blocks
        self.labels = self.output_labels
        super().__init__(ConvTrainNetwork(x_nodes, train_blocks, self.nodes, output_labels=self.labels))
        self.input_nodes = x_nodes
        self.main_nodes = train_blocks
        self.train_nodes = train_blocks + conv_blocks

    def grow_cart(self, new_bond=None, new_convolution_bond=None):
        """Append one column to a (trained) conv-TT (reference tensor/layers.py:892-947; image_convolution_growing_MNIST.py:88).

        The old last patch / pixel cores get a new right bond by broadcasting (every slice equal); the new patch core is
        ``1/new_bond`` on the bias patch and zero elsewhere, so with the new pixel core on the bias pixel the function is
        unchanged up to that pixel core's scale; the new pixel core is a fresh unit-norm random draw.  New tensors are created on
        the CPU, like the reference's: call ``.cuda()`` on the layer afterwards.  The network object is rebuilt (every cached
        environment belongs to the old topology); its engine settings carry over."""
        n = self.num_carriages
        Q, T = self.num_patches, self.patch_pixels
        new_bond = self.bond_dim if new_bond is None else new_bond
        CB = self.convolution_bond if new_convolution_bond is None else new_convolution_bond
        rb, cb = f"r{n + 1}", f"CB{n + 1}"
        x_new = TensorNode((1, Q, T), ["s", "patches", "patch_pixels"], name=f"X{n + 1}")
        core = torch.zeros((new_bond, 1, Q, 1))
        core[:, :, -1] = 1.0 / new_bond
        A_new = TensorNode(core, [rb, f"c{n + 1}", "patches", f"r{n + 2}"], l=rb, r=f"r{n + 2}", name=f"A{n + 1}")
        x_new.connect(A_new, "patches")
        if CB > 0:
            # the new last pixel core has no right bond (the reference's size expression reduces to this)
            C_new = TensorNode((CB if n != 1 else 1, T, 1), [cb, "patch_pixels", f"CB{n + 2}"], l=cb, r=f"CB{n + 2}", name=f"C{n + 1}")
        else:
            C_new = TensorNode((T,), ["patch_pixels"], name=f"C{n + 1}")
        x_new.connect(C_new, "patch_pixels")
        self.x_nodes.append(x_new)

        A_last, C_last = self.train_blocks[-1], self.conv_blocks[-1]
        A_last.expand_labels(A_last.dim_labels + [rb], tuple(A_last.shape) + (new_bond,))
        A_new.connect(A_last, rb)
        A_new.squeeze()
        self.train_blocks.append(A_new)
        if CB > 0:
            C_last.expand_labels(C_last.dim_labels + [cb], tuple(C_last.shape) + (CB,))
            C_last.connect(C_new, cb)
        C_new.squeeze()
        self.conv_blocks.append(C_new)
        self.num_carriages = n + 1

        from .conv import ConvTrainNetwork
        old = self.tensor_network
        self.tensor_network = ConvTrainNetwork(self.x_nodes, self.train_blocks, old.train_nodes + [C_new, A_new], output_labels=self.labels)
        for attr in ("chunk_rows", "dense_chunk_bytes", "process_group", "shard_offset", "shard_total", "gram_mode", "solve_mode"):
            setattr(self.tensor_network, attr, getattr(old, attr))
