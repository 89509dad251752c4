"""Packed, memory-mapped training set (input pipeline; opt-in with DALM_B200_PACKED_LOADER=1).

The reference tokenises with `datasets.map` and lets `DataLoader` + `default_data_collator` turn B python-list rows into int64
tensors per step (train_rage2e.py:306-334). Every feature of this path has a fixed width after padding / truncation (query 50,
passage 128, generator 256 ids + masks, one length scalar), so the tokenised set is ONE int32 matrix [N, W]: written once next
to the datasets cache (`.npy`, opened with mmap), a batch is a single fancy-indexed gather `[B, W]` from it and the step's
tensors are column slices of that block - no per-row python objects, no list -> tensor conversion, 4 bytes per token on disk
and in the page cache. The sampler / sharding / collation contract is unchanged: `PackedDataset` is a map-style dataset with a
batched `__getitems__`, its `collate` returns exactly what `loop.collate` returns for the same rows (tests/test_host_logic.py),
so the shuffled `DataLoader`, `_BatchSamplerShard` and the pinned H2D copy of the step work on it as they do on the HF dataset.

Measured through `train_e2e` the host side of a cfg-3 step is already hidden behind the asynchronous graph launch (DESIGN §2
f3), so this changes host CPU time per batch, not samples/s; it is off by default to keep the reference's pipeline as the
validated path.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def enabled() -> bool:
    return os.environ.get("DALM_B200_PACKED_LOADER", "0") == "1"


class PackedDataset(torch.utils.data.Dataset):
    """int32 [N, W] matrix + column table [(name, offset, width)]; width 0 marks a scalar feature (stored in one column)"""

    def __init__(self, matrix: np.ndarray, columns: Sequence[Tuple[str, int, int]]):
        self.matrix, self.columns = matrix, [(str(n), int(o), int(w)) for n, o, w in columns]

    # ---- construction ------------------------------------------------------------------------------------------------
    @staticmethod
    def layout(example: Dict[str, Any]) -> List[Tuple[str, int, int]]:
        cols, off = [], 0
        for k, v in example.items():
            w = len(v) if isinstance(v, (list, tuple)) else 0
            cols.append((k, off, w))
            off += max(w, 1)
        return cols

    @classmethod
    def from_rows(cls, rows, path: Optional[str] = None) -> "PackedDataset":
        """rows: an indexable of feature dicts (an HF dataset after tokenisation). Every list feature must have the same length
        in every row (the reference pads / truncates to max_length), every other feature must be an int. path: where to keep the
        matrix (`<path>.npy` + `<path>.json`); None = in memory."""
        n = len(rows)
        if n == 0:
            raise ValueError("PackedDataset: empty dataset")
        cols = cls.layout(rows[0])
        W = sum(max(w, 1) for _, _, w in cols)
        if path is not None and os.path.exists(path + ".npy") and os.path.exists(path + ".json"):
            meta = json.load(open(path + ".json"))
            if meta.get("rows") == n and [tuple(c) for c in meta.get("columns", [])] == cols:
                return cls(np.load(path + ".npy", mmap_mode="r"), cols)
        if path is not None:
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            tmp = f"{path}.{os.getpid()}.tmp.npy"
            mat = np.lib.format.open_memmap(tmp, mode="w+", dtype=np.int32, shape=(n, W))
        else:
            mat = np.empty((n, W), dtype=np.int32)
        if hasattr(rows, "column_names") and hasattr(rows, "__getitem__"):
            for name, off, w in cols:                            # HF dataset: one arrow column at a time
                col = np.asarray(rows[name])
                if w == 0:
                    mat[:, off] = col.astype(np.int64)
                else:
                    if col.ndim != 2 or col.shape[1] != w:
                        raise ValueError(f"PackedDataset: feature {name!r} is not fixed-width ({w}) in every row")
                    mat[:, off:off + w] = col
        else:
            for i in range(n):
                r = rows[i]
                for name, off, w in cols:
                    v = r[name]
                    if w == 0:
                        mat[i, off] = int(v)
                    else:
                        if len(v) != w:
                            raise ValueError(f"PackedDataset: feature {name!r} of row {i} has {len(v)} entries, expected {w}")
                        mat[i, off:off + w] = v
        if path is not None:
            mat.flush()
            del mat
            os.replace(tmp, path + ".npy")                       # atomic: concurrent ranks either see the whole file or none
            with open(path + ".json", "w") as f:
                json.dump({"rows": n, "columns": cols}, f)
            mat = np.load(path + ".npy", mmap_mode="r")
        return cls(mat, cols)

    @staticmethod
    def cache_path(rows, cache_root: Optional[str]) -> Optional[str]:
        """a stable file name for a tokenised HF dataset (its fingerprint changes with the data, the tokenizer and the lengths)"""
        fp = getattr(rows, "_fingerprint", None)
        if cache_root is None or fp is None:
            return None
        return os.path.join(cache_root, "packed_" + hashlib.sha1(str(fp).encode()).hexdigest()[:16])

    # ---- dataset protocol ------------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(self.matrix.shape[0])

    def __getitem__(self, i: int) -> np.ndarray:
        return np.asarray(self.matrix[int(i)])

    def __getitems__(self, indices: Sequence[int]) -> np.ndarray:
        """the whole batch with ONE gather (torch's fetcher calls this when it exists)"""
        return np.ascontiguousarray(self.matrix[np.asarray(indices, dtype=np.int64)])

    def collate(self, block) -> Dict[str, torch.Tensor]:
        """[B, W] int32 block (or a list of rows from the per-item path) -> the dict `loop.collate` builds: int64 tensors
        [B, width], scalars as [B]"""
        if isinstance(block, list):
            block = np.stack(block, 0)
        t = torch.from_numpy(block).to(torch.int64)
        return {name: (t[:, off] if w == 0 else t[:, off:off + w]).contiguous() for name, off, w in self.columns}
