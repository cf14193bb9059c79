"""Labelled tensor node: the graph vocabulary the layer constructors speak.

Mirror of the reference's ``TensorNode`` (tensor/node.py:6-282) as far as the sweep path and the
layer constructors use it: a tensor, one label per dimension, which labels are left/right bonds,
and which neighbour each label leads to.  The B200 engine (``network.py``) only *reads* this
graph to recognise the chain; it does not contract through it, so the pairwise label-driven
einsum of the reference (node.py:28-74) survives here only as a small convenience for callers
that build custom pieces on the host.
"""
import string
from collections import defaultdict

import torch


class TensorNode:
    def __init__(self, tensor_or_shape, dim_labels, l=None, r=None, name=None, dtype=None):
        if isinstance(tensor_or_shape, (tuple, list)):
            # random core of unit Frobenius norm (reference node.py:9-11); same RNG stream as torch.randn
            t = torch.randn(tensor_or_shape, dtype=dtype)
            self.tensor = t / torch.norm(t)
        else:
            self.tensor = tensor_or_shape
        self.dim_labels = list(dim_labels)
        self.left_labels = self._as_list(l)
        self.right_labels = self._as_list(r)
        self.name = name or ""
        self.connections = {}
        self.connection_priority = defaultdict(float)
        self.contracted = set()

    @staticmethod
    def _as_list(v):
        if v is None:
            return []
        return [v] if isinstance(v, str) else list(v)

    # -- graph ---------------------------------------------------------------------------
    def connect(self, other, labels, priority=float("-inf")):
        for lab in ([labels] if isinstance(labels, str) else list(labels)):
            for a, b in ((self, other), (other, self)):
                a.connection_priority[lab] = max(a.connection_priority[lab], priority) if lab in a.connections else priority
                a.connections[lab] = b

    def reset_connections(self):
        self.connections = {}
        self.connection_priority = defaultdict(float)
        self.contracted = set()

    def is_horizontal_bond(self, label):
        return label in self.left_labels or label in self.right_labels

    def get_connecting_labels(self, other, horizontal=True):
        def direct(a, b):
            return {lab for lab, n in a.connections.items()
                    if n is b and (horizontal or not a.is_horizontal_bond(lab))}
        mine = self.contracted or {self}
        theirs = other.contracted | {other}
        out = set()
        for a in mine:
            for b in theirs:
                out |= direct(a, b) | direct(b, a)
        return list(out)

    # -- tensor views ----------------------------------------------------------------------
    @property
    def shape(self):
        return self.tensor.shape

    def dim_size(self, label):
        return self.tensor.shape[self.dim_labels.index(label)]

    def set_tensor(self, tensor):
        self.tensor = tensor
        return self

    def cuda(self):
        self.tensor = self.tensor.cuda()
        return self

    def to(self, device=None, dtype=None):
        self.tensor = self.tensor.to(device=device, dtype=dtype)
        return self

    def copy(self):
        return TensorNode(self.tensor, list(self.dim_labels), l=list(self.left_labels), r=list(self.right_labels),
                          name=self.name + "_c")

    def squeeze(self, exclude=()):
        """Drop size-1 legs that lead nowhere and are not protected (reference node.py:135-147)."""
        drop = [i for i, (s, lab) in enumerate(zip(self.shape, self.dim_labels))
                if s <= 1 and lab not in exclude and lab not in self.connections]
        if drop:
            gone = {self.dim_labels[i] for i in drop}
            self.tensor = self.tensor.squeeze(*drop)
            self.dim_labels = [lab for lab in self.dim_labels if lab not in gone]
            self.left_labels = [lab for lab in self.left_labels if lab not in gone]
            self.right_labels = [lab for lab in self.right_labels if lab not in gone]
        return self

    def permute(self, *labels):
        self.tensor = self.tensor.permute(*[self.dim_labels.index(lab) for lab in labels])
        self.dim_labels = list(labels)
        return self

    def expand_labels(self, labels, size):
        """Append the missing labels as new trailing legs and broadcast the listed legs to ``size`` (a stride-0 view, as in
        the reference, node.py:243-253; the first update of the node replaces it by a dense tensor)."""
        labels = list(labels)
        for lab in labels:
            if lab not in self.dim_labels:
                self.tensor = self.tensor.unsqueeze(-1)
                self.dim_labels = self.dim_labels + [lab]
        self.tensor = self.tensor.expand(*[size[labels.index(lab)] if lab in labels else -1 for lab in self.dim_labels])
        return self

    def permute_first(self, *labels, expand=True):
        rest = [lab for lab in self.dim_labels if lab not in labels]
        order = [lab for lab in list(labels) + rest if expand or lab in self.dim_labels]
        present = [lab for lab in order if lab in self.dim_labels]
        if present:
            self.tensor = self.tensor.permute(*[self.dim_labels.index(lab) for lab in present])
        if expand:
            for pos, lab in enumerate(order):
                if lab not in self.dim_labels:
                    self.tensor = self.tensor.unsqueeze(pos)
        self.dim_labels = order
        return self

    def permute_last(self, *labels):
        rest = [lab for lab in self.dim_labels if lab not in labels]
        order = rest + list(labels)
        present = [lab for lab in order if lab in self.dim_labels]
        self.tensor = self.tensor.permute(*[self.dim_labels.index(lab) for lab in present])
        for pos, lab in enumerate(order):
            if lab not in self.dim_labels:
                self.tensor = self.tensor.unsqueeze(pos)
        self.dim_labels = order
        return self

    def sum_labels(self, labels):
        labels = [labels] if isinstance(labels, str) else labels
        return self.tensor.sum([self.dim_labels.index(lab) for lab in labels if lab in self.dim_labels])

    def contract_with(self, other, contract_labels=None):
        """Host-side pairwise contraction by label (small tensors only; not on the sweep path)."""
        if self is other:
            return self
        if contract_labels is None:
            contract_labels = self.get_connecting_labels(other)
        contract_labels = [contract_labels] if isinstance(contract_labels, str) else list(contract_labels)
        labels = list(dict.fromkeys(self.dim_labels + other.dim_labels))
        letter = {lab: string.ascii_letters[i] for i, lab in enumerate(labels)}
        keep = [lab for lab in labels if lab not in contract_labels]
        spec = "".join(letter[x] for x in self.dim_labels) + "," + "".join(letter[x] for x in other.dim_labels) + "->" + "".join(letter[x] for x in keep)
        out = TensorNode(torch.einsum(spec, self.tensor, other.tensor), keep,
                         l=[x for x in self.left_labels + other.left_labels if x not in contract_labels],
                         r=[x for x in self.right_labels + other.right_labels if x not in contract_labels],
                         name=f"<{self.name}-{','.join(contract_labels)}-{other.name}>")
        out.contracted = (self.contracted or {self}) | (other.contracted or {other})
        for src in (self, other):
            for lab, n in src.connections.items():
                if n not in out.contracted:
                    out.connection_priority[lab] = max(out.connection_priority[lab], src.connection_priority[lab]) if lab in out.connections else src.connection_priority[lab]
                    out.connections[lab] = n
        return out

    def update_node(self, step, lr=1.0, adaptive_step=False, min_norm=None, max_norm=None):
        """theta <- theta + lr*step on the device (reference node.py:178-203)."""
        from .. import ops
        new = self.tensor.detach().clone().contiguous()
        ops.update_node(new.view(-1), step.contiguous().view(-1), lr=lr, adaptive_step=adaptive_step, max_norm=max_norm)
        self.tensor = new
        return self

    def __repr__(self):
        return f"TensorNode(name={self.name}, shape={tuple(self.shape)}, labels={self.dim_labels})"
