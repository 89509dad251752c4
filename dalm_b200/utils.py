"""dalm.utils equivalents (reference dalm/utils.py:8-35)."""
from __future__ import annotations

import os
from typing import Any

import torch


def load_dataset(dataset_or_path: Any):
    """what the trainers and evaluators accept as data (reference :8-19): an in-memory `datasets.Dataset` (returned as is), a
    directory written by `Dataset.save_to_disk`, or anything else handed to the csv loader (its "train" split)"""
    import datasets

    if isinstance(dataset_or_path, datasets.Dataset):
        return dataset_or_path
    source = os.fspath(dataset_or_path)
    if os.path.isdir(source):
        return datasets.load_from_disk(source)
    return datasets.load_dataset("csv", data_files=source)["train"]


def eos_mask(mask: torch.Tensor, padding: str = "left") -> torch.Tensor:
    """one-hot [B, L] selecting the last real token of each row (reference :22-35): the last column, or — with right padding —
    column (number of ones - 1). Pure index arithmetic; same dtype as `mask`."""
    if padding == "right":
        last = (mask.sum(dim=1, keepdim=True) - 1) % mask.size(1)      # an all-zero row selects the last column (index -1 there)
    else:
        last = torch.full((mask.size(0), 1), mask.size(1) - 1, dtype=torch.int64, device=mask.device)
    return torch.zeros_like(mask).scatter_(1, last.to(torch.int64), 1)
