"""reference dalm/eval/eval_results.py:1-8"""
from pydantic import BaseModel


class EvalResults(BaseModel):
    total_examples: int
    recall: float
    precision: float
    hit_rate: float
