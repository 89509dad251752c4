"""`dalm eval-rag` — reference dalm/eval/eval_rag.py:167-290. The retriever half (passage sweep, top-k, recall / precision /
hit-rate through `rag_model.retrieval_forward`) is built; the generator half (`model.generate` + exact match, :118-165,
:258-283) needs an autoregressive KV-cache decode path, which is a different workload from the training step this build
covers: `evaluate_generator=True` raises NotImplementedError instead of silently skipping it."""
from __future__ import annotations

import logging
from typing import Any, Final, Literal, Optional

import torch
from torch.utils.data import DataLoader

from ..models.rag_e2e_base_model import AutoModelForRagE2E, inference_only
from ..utils import load_dataset
from .eval_results import EvalResults
from .utils import (calc_eval_results, construct_search_index, evaluate_retriever_on_batch, get_passage_embeddings,
                    mixed_collate_fn, preprocess_dataset, print_eval_results)

logger = logging.getLogger(__name__)


def evaluate_rag(
    dataset_or_path: Any,
    retriever_name_or_path: str,
    generator_name_or_path: str,
    retriever_peft_model_path: Optional[str],
    generator_peft_model_path: Optional[str],
    passage_column_name: str,
    query_column_name: str,
    answer_column_name: str,
    embed_dim: int,
    max_length: int,
    test_batch_size: int = 8,
    query_batch_size: int = 16,
    device: str = "cuda",
    torch_dtype: Literal["float16", "bfloat16"] = "float16",
    top_k: int = 10,
    evaluate_generator: bool = True,
    retriever_is_autoregressive: bool = False,
) -> EvalResults:
    if evaluate_generator:
        raise NotImplementedError("eval-rag's generator evaluation (generate + exact match, reference eval_rag.py:118-165) needs "
                                  "an autoregressive decode path that dalm_b200 does not build; pass evaluate_generator=False "
                                  "(--no-evaluate-generator) for the retriever metrics")
    if not str(device).startswith("cuda"):
        raise RuntimeError("dalm_b200 evaluates on a CUDA (sm_100a) device only: there is no CPU path")
    test_dataset = load_dataset(dataset_or_path)
    selected_torch_dtype: Final[torch.dtype] = torch.float16 if torch_dtype == "float16" else torch.bfloat16
    with inference_only():
        rag_model = AutoModelForRagE2E(retriever_name_or_path, generator_name_or_path,
                                       retriever_is_autoregressive=retriever_is_autoregressive)
    rag_model.eval()
    processed = preprocess_dataset(test_dataset, rag_model.retriever_tokenizer, query_column_name, passage_column_name, max_length)
    rag_model.attach_pre_trained_peft_layers(retriever_peft_model_path, generator_peft_model_path, device)
    dev = str(rag_model.retriever_model.dev)
    unique_passage_dataset, passage_embeddings = get_passage_embeddings(processed, passage_column_name, rag_model.retrieval_forward,
                                                                       dev, embed_dim, selected_torch_dtype, test_batch_size)
    id_to_passage = {i: p[passage_column_name] for i, p in enumerate(unique_passage_dataset)}
    index = construct_search_index(embed_dim, len(passage_embeddings), passage_embeddings)
    batch_precision, batch_recall, total_hit = [], [], 0
    loader = DataLoader(processed, batch_size=test_batch_size, shuffle=True, collate_fn=mixed_collate_fn)
    for batch in loader:
        p_, r_, h_, _ = evaluate_retriever_on_batch(batch, passage_column_name, rag_model.retrieval_forward, index,
                                                    selected_torch_dtype, dev, top_k, id_to_passage)
        batch_precision.extend(p_)
        batch_recall.extend(r_)
        total_hit += h_
    results = calc_eval_results(len(processed), batch_precision, batch_recall, total_hit)
    print_eval_results(results)
    return results
