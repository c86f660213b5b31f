"""On-disk formats of the reference around the path: checkpoints in the layout of
``utils/model_utils.save_model / load_model`` (:19-38) and the two result files ``test.py`` writes (:40-44), so that
models and results move between the reference and the drop-in in both directions."""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch


def save_model(model, model_dir: str, current_epoch: int, last_best_epoch: Optional[int] = None, name: str = "training") -> str:
    """utils/model_utils.py:19-31: ``{name}_model_epoch{E}.pth`` holding {'model_state_dict', 'epoch'}; the file of
    the previous best epoch is removed.  Row partitioned models: call ``model.complete_attention()`` on every rank
    first (``state_dict()`` refuses otherwise)."""
    os.makedirs(model_dir, exist_ok=True)
    path = os.path.join(model_dir, f"{name}_model_epoch{current_epoch}.pth")
    torch.save({"model_state_dict": model.state_dict(), "epoch": current_epoch}, path)
    if last_best_epoch is not None and current_epoch != last_best_epoch:
        old = os.path.join(model_dir, f"{name}_model_epoch{last_best_epoch}.pth")
        if os.path.exists(old):
            os.remove(old)
    return path


def load_model(model, model_path: str):
    """utils/model_utils.py:34-38 (checkpoints written by the reference load unchanged: same state-dict keys, the
    sparse ``A_in`` entry included)."""
    checkpoint = torch.load(model_path, map_location=torch.device("cpu"), weights_only=False)
    model.load_state_dict(checkpoint["model_state_dict"])
    model.eval()
    return model


def metrics_line(elapsed_s: float, metrics: Dict[str, float]) -> str:
    """The line test.py:30-31 formats."""
    return ("Running test: Total Time {:.1f}s | Accuracy [{:.4f}], Precision [{:.4f}], Recall [{:.4f}], F1 [{:.4f}]"
            .format(elapsed_s, metrics["accuracy"], metrics["precision"], metrics["recall"], metrics["f1"]))


def write_test_results(save_dir: str, elapsed_s: float, metrics: Dict[str, float], prediction_scores,
                       reference_paths: bool = True) -> Dict[str, str]:
    """test.py:40-44: ``test_results.tsv`` (one column ``metrics`` with the formatted line) and the raw
    ``prediction_scores.npy``.  ``reference_paths``: reproduce upstream's file names exactly -- it concatenates
    ``save_dir + 'prediction_scores.npy'`` without a separator (test.py:44)."""
    os.makedirs(save_dir, exist_ok=True)
    tsv = save_dir + "/test_results.tsv"
    with open(tsv, "w") as fh:                       # pandas.DataFrame([{metrics: line}]).to_csv(sep='\t', index=False)
        fh.write("metrics\n" + metrics_line(elapsed_s, metrics) + "\n")
    npy = (save_dir + "prediction_scores.npy") if reference_paths else os.path.join(save_dir, "prediction_scores.npy")
    scores = prediction_scores.detach().cpu().numpy() if isinstance(prediction_scores, torch.Tensor) else np.asarray(prediction_scores)
    np.save(npy, scores)
    return {"tsv": tsv, "npy": npy}
