"""More accumulating_swipe recordings of the reference, for the keyword semantics of SURVEY.md Appendix D that the first eight
fixtures do not exercise: direction='r2l' with eps_per_node (the growing-TT call shape), adaptive_step + max_norm + lr < 1.
Build container only:    python tests/golden/make_golden_keywords.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (stubs matplotlib, puts /root/reference on sys.path)
import torch  # noqa: E402
from tensor.layers import TensorTrainLayer  # noqa: E402
from tensor.bregman import SquareBregFunction  # noqa: E402


def main():
    rng = np.random.default_rng(99)
    N, F = 230, 3
    X = rng.uniform(-1, 1, size=(N, F))
    Xb = torch.tensor(np.concatenate([X, np.ones((N, 1))], 1))
    y = torch.tensor(mg.teacher(X, 12))
    # one right-to-left pass with one epsilon per node (indexed by position in the node order, network.py:427-431)
    layer = TensorTrainLayer(4, 3, F + 1, output_shape=1, constrict_bond=False, perturb=True, seed=21)
    kw = dict(batch_size=90, num_swipes=1, lr=1.0, method="ridge_cholesky", eps=[1.0, 0.3, 0.1, 0.03], skip_second=True,
              direction="r2l", eps_per_node=True)
    rec = mg.record_swipe(layer, Xb, y, SquareBregFunction(), **kw)
    mg.save("tt_r2l_pernode", rec, Xb, y, dict(kind="tt", n=4, r=3, f=F + 1, C=1, loss="square", perturb=True, **kw))
    # step control of TensorNode.update_node (node.py:178-203): adaptive shrink, max-norm projection, lr < 1
    layer = TensorTrainLayer(3, 3, F + 1, output_shape=1, constrict_bond=False, perturb=False, seed=22)
    kw = dict(batch_size=-1, num_swipes=1, lr=0.7, method="ridge_cholesky", eps=0.5, eps_decay=0.5, adaptive_step=True, max_norm=0.8)
    rec = mg.record_swipe(layer, Xb, y, SquareBregFunction(), **kw)
    mg.save("tt_adaptive_maxnorm", rec, Xb, y, dict(kind="tt", n=3, r=3, f=F + 1, C=1, loss="square", perturb=False, **kw))


if __name__ == "__main__":
    main()
