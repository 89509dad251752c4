"""CPU: dalm_b200's batch builders produce BIT-EXACT token ids / lengths vs the reference's builders (golden fixture
generated from the reference by oracle/make_golden.py) — north_star: "bit-exact token indices"."""
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fixture():
    from transformers import AutoTokenizer
    with open(os.path.join(GOLD, "preprocess.json")) as f:
        gold = json.load(f)
    rt = AutoTokenizer.from_pretrained(os.path.join(GOLD, "tok_bert"))
    gt = AutoTokenizer.from_pretrained(os.path.join(GOLD, "tok_llama"))
    gt.pad_token = gt.eos_token
    gt.add_eos_token = True
    ex = {k: [r[k] for r in gold["rows"]] for k in ("Abstract", "Question", "Answer")}
    return gold, rt, gt, ex


def test_e2e_builder_bit_exact(fixture):
    from dalm_b200.training.utils.rag_e2e_dataloader_utils import preprocess_dataset
    gold, rt, gt, ex = fixture
    got = preprocess_dataset(ex, retriever_tokenizer=rt, generator_tokenizer=gt, query_column_name="Question",
                             passage_column_name="Abstract", answer_column_name="Answer", query_max_len=50,
                             passage_max_len=128, generator_max_len=256)
    assert set(got) == set(gold["e2e"])
    for k, v in gold["e2e"].items():
        assert [list(x) if isinstance(x, (list, tuple)) else x for x in got[k]] == v, k
    assert all(len(x) == 50 for x in got["retriever_query_input_ids"])
    assert all(len(x) == 128 for x in got["retriever_passage_input_ids"])
    assert all(len(x) == 256 for x in got["generator_input_input_ids"])
    assert all(q >= 1 for q in got["query_passage_input_len"])


def test_retriever_builder_bit_exact(fixture):
    from dalm_b200.training.utils.retriever_only_dataloader_utils import preprocess_dataset
    gold, rt, gt, ex = fixture
    got = preprocess_dataset(ex, rt, query_column_name="Question", passage_column_name="Abstract", query_max_len=50,
                             passage_max_len=128)
    assert set(got) == set(gold["retriever"])
    for k, v in gold["retriever"].items():
        assert [list(x) for x in got[k]] == v, k


def test_ragged_columns_rejected(fixture):
    from dalm_b200.training.utils.rag_e2e_dataloader_utils import preprocess_dataset
    gold, rt, gt, ex = fixture
    bad = dict(ex, Answer=ex["Answer"][:-1])
    with pytest.raises(ValueError):
        preprocess_dataset(bad, rt, gt, "Question", "Abstract", "Answer", 50, 128, 256)
