"""On-disk formats either side of the sweep path, as the reference's scripts read and write them.

* datasets: one ``torch.save`` dict per dataset with the float64 splits ``X_train, y_train, X_val, y_val, X_test, y_test``
  (reference prep_file/prepare_higgs.py:42-52 and siblings; ``data/<name>_tensor.pt``);
* grid-search results: one CSV row per (dataset, N, r, seed) with the run settings appended as constant columns
  (reference train_grid_search_ablation_tt.py:89-99).
"""
import torch

DATASET_KEYS = ("X_train", "y_train", "X_val", "y_val", "X_test", "y_test")
RESULT_COLUMNS = ["dataset", "N", "r", "lin_dim", "val_rmse", "val_r2", "val_accuracy", "num_params", "converged_epoch", "seed"]
RUN_COLUMNS = ["num_swipes", "eps_start", "eps_decay", "early_stopping", "model_type"]


def save_tensor_dataset(path, **splits):
    missing = [k for k in DATASET_KEYS if k not in splits]
    if missing:
        raise ValueError(f"missing splits: {missing}")
    torch.save({k: torch.as_tensor(splits[k], dtype=torch.float64).cpu() for k in DATASET_KEYS}, path)


def load_tensor_dataset(path, device="cuda"):
    """The six splits of a ``*_tensor.pt`` file, moved to ``device`` (the B200 path has no CPU arithmetic)."""
    d = torch.load(path, map_location="cpu")
    missing = [k for k in DATASET_KEYS if k not in d]
    if missing:
        raise ValueError(f"{path}: not a dataset file of the reference (missing {missing})")
    return {k: d[k].to(device=device, dtype=torch.float64) for k in DATASET_KEYS}


def results_frame(rows, num_swipes, eps_start, eps_decay, early_stopping, model_type):
    """``rows``: tuples in RESULT_COLUMNS order.  Returns the DataFrame the reference writes with ``to_csv(index=False)``."""
    import pandas as pd
    df = pd.DataFrame(list(rows), columns=RESULT_COLUMNS)
    for k, v in zip(RUN_COLUMNS, (num_swipes, eps_start, eps_decay, early_stopping, model_type)):
        df[k] = v
    return df
