"""Batch builder of the retriever-only trainer (reference dalm/training/utils/retriever_only_dataloader_utils.py:8-27):
`#query# ` / `#passage# ` prefixes, pad/truncate to fixed lengths, keys `query_*` / `passage_*`."""
from __future__ import annotations

from typing import Any, Dict

from .rag_e2e_dataloader_utils import P_TAG, Q_TAG, _tag


def preprocess_dataset(examples: Any, tokenizer: Any, query_column_name: str, passage_column_name: str,
                       query_max_len: int, passage_max_len: int) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    spec = (("query_", _tag(Q_TAG, examples[query_column_name]), query_max_len),
            ("passage_", _tag(P_TAG, examples[passage_column_name]), passage_max_len))
    for prefix, texts, max_len in spec:
        enc = tokenizer(texts, padding="max_length", max_length=max_len, truncation=True)
        out.update({prefix + k: v for k, v in enc.items()})
    return out
