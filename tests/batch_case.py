"""Shared body of the minibatch-estimator (TensorTrainBatchRegressor) parity test: tests/golden/make_golden_batch.py."""
import os

import numpy as np

import golden_util as gu

CASES = {"unique": dict(swipe_method="batch_unique", num_swipes=3, N=3, r=3, batch_size=64, eps_start=2.0, eps_end=0.5),
         "same": dict(swipe_method="batch_same", num_swipes=2, N=3, r=3, batch_size=96, eps_start=2.0, eps_end=0.5, perturb=False),
         "block": dict(swipe_method="batch_block", num_swipes=2, N=4, r=2, batch_size=80, eps_start=1.0, eps_end=1.0)}


def run(tag, device):
    """(max trajectory error, prediction error, max relative core error) against the reference recording."""
    from tensornetworksfork_b200.tensor.module import TensorTrainBatchRegressor
    z = np.load(os.path.join(gu.GOLDEN_DIR, "batch_tt.npz"))
    est = TensorTrainBatchRegressor(device=device, seed=5, **CASES[tag])
    est.fit(z["X"], z["y"])
    traj = np.array([[t["epoch"], t["val_rmse"]] for t in est.trajectory])
    assert traj.shape == z[f"{tag}_traj"].shape, (traj.shape, z[f"{tag}_traj"].shape)
    assert np.array_equal(traj[:, 0], z[f"{tag}_traj"][:, 0])
    traj_err = float(np.max(np.abs(traj[:, 1] - z[f"{tag}_traj"][:, 1])))
    pred_err = gu.relerr(est.predict(z["X"]), z[f"{tag}_pred"])
    core_err = max(gu.relerr(nd.tensor.cpu().numpy(), z[f"{tag}_core_{i}"]) for i, nd in enumerate(est._model.tensor_network.train_nodes))
    return traj_err, pred_err, core_err
