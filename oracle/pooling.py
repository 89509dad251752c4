"""CPU restatement of mean_pooling / normalize / eos_mask.  TEST INFRASTRUCTURE ONLY.

reference: dalm/models/rag_e2e_base_model.py:96-97,108-111 (= retriever_only_base_model.py:60-68); dalm/utils.py:22-35
"""
from __future__ import annotations

import torch


def mean_pooling(token_embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    m = attention_mask.unsqueeze(-1).expand(token_embeddings.size()).float()
    return torch.sum(token_embeddings * m, 1) / torch.clamp(m.sum(1), min=1e-9)


def normalize(x: torch.Tensor) -> torch.Tensor:
    # torch.nn.functional.normalize(p=2, dim=1): x / max(||x||_2, 1e-12)
    return x / torch.clamp(x.norm(dim=1, keepdim=True), min=1e-12)


def eos_mask(mask: torch.Tensor, padding: str = "left") -> torch.Tensor:
    new = torch.zeros_like(mask)
    if padding == "right":
        cnt = mask.sum(dim=1)
        new[torch.arange(mask.size(0)), cnt - 1] = 1
    else:
        new[:, -1] = 1
    return new
