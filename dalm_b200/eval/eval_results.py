"""The record `evaluate_retriever` / `evaluate_rag` return — the four fields of the reference's dalm/eval/eval_results.py:1-8
(a pydantic model there too, so `.dict()` / `.json()` keep working for callers), documented, plus the log lines the
evaluators print."""
from __future__ import annotations

from typing import List

from pydantic import BaseModel, Field


class EvalResults(BaseModel):
    total_examples: int = Field(description="rows of the evaluation set (the denominator of every mean below)")
    recall: float = Field(description="mean over rows of |retrieved ∩ {gold passage}| / 1")
    precision: float = Field(description="mean over rows of |retrieved ∩ {gold passage}| / |distinct retrieved passages|")
    hit_rate: float = Field(description="fraction of rows whose gold passage is among the top-k")

    def log_lines(self) -> List[str]:
        """the block `print_eval_results` logs (reference dalm/eval/utils.py:288-295)"""
        return ["Retriever results:", f"Recall: {self.recall}", f"Precision: {self.precision}", f"Hit Rate: {self.hit_rate}", "*" * 13]
