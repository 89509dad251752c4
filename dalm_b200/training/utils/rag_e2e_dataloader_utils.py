"""Batch builder of the RAG-e2e trainer: text -> fixed-length token-id lists.

Behaviour-identical to the reference's dalm/training/utils/rag_e2e_dataloader_utils.py:7-68, including its quirks
(SURVEY §8a): the generator text is built from the ALREADY-PREFIXED query / passage strings (so "#query# #query# ..."),
and `query_passage_input_len` counts the untruncated, unpadded tokens of the "...#answer#" prompt (BOS/EOS included as
the tokenizer emits them). Token ids must be bit-exact: tests/test_preprocess.py checks this against outputs of the
reference function committed under tests/golden/.
"""
from __future__ import annotations

from typing import Any, Dict, List

Q_TAG, P_TAG, A_TAG = "#query#", "#passage#", "#answer#"


def _tag(tag: str, texts: List[str]) -> List[str]:
    return [f"{tag} {t}" for t in texts]


def preprocess_dataset(
    examples: Any,
    retriever_tokenizer: Any,
    generator_tokenizer: Any,
    query_column_name: str,
    passage_column_name: str,
    answer_column_name: str,
    query_max_len: int,
    passage_max_len: int,
    generator_max_len: int,
) -> Dict[str, Any]:
    raw_q, raw_p, answers = examples[query_column_name], examples[passage_column_name], examples[answer_column_name]
    if not (len(raw_q) == len(raw_p) == len(answers)):
        raise ValueError("query / passage / answer columns differ in length")       # zip(strict=True) in the reference
    queries, passages = _tag(Q_TAG, raw_q), _tag(P_TAG, raw_p)

    fixed = dict(padding="max_length", truncation=True)
    batch: Dict[str, Any] = {}
    for prefix, texts, max_len in (("retriever_query_", queries, query_max_len),
                                   ("retriever_passage_", passages, passage_max_len)):
        for key, val in retriever_tokenizer(texts, max_length=max_len, **fixed).items():
            batch[prefix + key] = val

    # prompt = tagged query + tagged passage + answer tag (the tags of `queries` / `passages` are repeated: reference :35-38)
    prompts = [f"{Q_TAG} {q} {P_TAG} {p} {A_TAG}" for q, p in zip(queries, passages)]
    full_text = [f"{pr} {a}" for pr, a in zip(prompts, answers)]
    for key, val in generator_tokenizer(full_text, max_length=generator_max_len, **fixed).items():
        batch["generator_input_" + key] = val

    prompt_ids = generator_tokenizer(prompts, padding=False)["input_ids"]
    batch["query_passage_input_len"] = [len(ids) for ids in prompt_ids]
    return batch
