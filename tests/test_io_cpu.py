"""Dataset (.pt dict) and result-CSV formats of the reference's scripts (prepare_higgs.py:42-52, train_grid_search_ablation_tt.py:89-99)."""
import numpy as np
import pandas as pd
import pytest
import torch

from tensornetworksfork_b200 import io as tio


def test_dataset_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    sp = {k: rng.normal(size=(7, 3) if k.startswith("X") else (7,)) for k in tio.DATASET_KEYS}
    path = tmp_path / "toy_tensor.pt"
    tio.save_tensor_dataset(path, **sp)
    raw = torch.load(path)
    assert set(raw) == set(tio.DATASET_KEYS) and all(v.dtype == torch.float64 for v in raw.values())
    got = tio.load_tensor_dataset(path, device="cpu")
    for k in tio.DATASET_KEYS:
        assert np.array_equal(got[k].numpy(), sp[k])
    torch.save({"X_train": raw["X_train"]}, tmp_path / "bad.pt")
    with pytest.raises(ValueError):
        tio.load_tensor_dataset(tmp_path / "bad.pt", device="cpu")


def test_results_csv_schema(tmp_path):
    rows = [("abalone", 3, 6, np.nan, 0.21, 0.55, np.nan, 1234, 7, 42), ("abalone", 3, 6, np.nan, 0.22, 0.54, np.nan, 1234, 9, 43)]
    df = tio.results_frame(rows, num_swipes=30, eps_start=1.0, eps_decay=0.5, early_stopping=5, model_type="tt")
    p = tmp_path / "abalone_ablation_results_tt.csv"
    df.to_csv(p, index=False)
    back = pd.read_csv(p)
    assert list(back.columns) == ["dataset", "N", "r", "lin_dim", "val_rmse", "val_r2", "val_accuracy", "num_params", "converged_epoch",
                                  "seed", "num_swipes", "eps_start", "eps_decay", "early_stopping", "model_type"]
    agg = back.groupby(["N", "r"]).agg({"val_rmse": "mean"}).reset_index()       # what the reference does next (:103)
    assert abs(agg["val_rmse"][0] - 0.215) < 1e-12
