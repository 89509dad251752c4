"""dalm.utils equivalents (reference dalm/utils.py:8-35)."""
from __future__ import annotations

import os
from typing import Any

import torch


def load_dataset(dataset_or_path: Any):
    """csv file -> datasets 'csv' loader; directory -> load_from_disk; Dataset -> passthrough (reference :8-19)"""
    import datasets

    if isinstance(dataset_or_path, datasets.Dataset):
        return dataset_or_path
    if os.path.isdir(dataset_or_path):
        return datasets.load_from_disk(dataset_or_path)
    return datasets.load_dataset("csv", data_files=dataset_or_path)["train"]


def eos_mask(mask: torch.Tensor, padding: str = "left") -> torch.Tensor:
    """one-hot mask selecting the last real token of each sequence (reference :22-35): with right padding that is
    position count-1, otherwise the last column. Pure index arithmetic."""
    picked = torch.zeros_like(mask)
    if padding == "right":
        last = mask.sum(dim=1) - 1
        picked[torch.arange(mask.size(0), device=mask.device), last] = 1
    else:
        picked[:, -1] = 1
    return picked
