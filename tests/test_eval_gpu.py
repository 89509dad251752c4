"""-m gpu: the evaluation path: exact inner-product top-k kernel against the float64 oracle (bit-exact indices),
the hnswlib-shaped index object, per-batch retriever metrics, and `evaluate_retriever` end to end incl. attaching trained adapters."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nq,N,D,K", [(1, 5, 32, 1), (8, 1000, 384, 10), (9, 20011, 1024, 10), (20, 3000, 64, 32),
                                      (3, 40, 128, 32), (16, 70000, 1024, 5)])
def test_topk_indices_bit_exact_vs_oracle(cuda_dev, nq, N, D, K):
    from dalm_b200 import ops
    from oracle import topk
    g = torch.Generator().manual_seed(nq * 7 + N)
    P = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(nq, D, generator=g), dim=1)
    scores, idx = ops.topk_ip(Q.to(cuda_dev), P.to(cuda_dev), K)
    labels, dist = topk.knn_query(P.numpy(), Q.numpy(), K)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), labels)            # integer work: bit-exact
    assert np.allclose(scores.cpu().numpy(), 1.0 - dist, atol=2e-6)
    assert (scores[:, :-1] >= scores[:, 1:]).all()                               # size-independent property: sorted


def test_topk_ties_padding_and_strided_rows(cuda_dev):
    from dalm_b200 import ops
    P = torch.zeros(10, 16)
    P[:, 0] = torch.tensor([1., 5., 5., 2., 5., 0., -1., 5., 3., 3.])
    Q = torch.zeros(2, 16); Q[0, 0] = 1.0; Q[1, 0] = -1.0
    s, i = ops.topk_ip(Q.to(cuda_dev), P.to(cuda_dev), 6)
    assert i[0].tolist() == [1, 2, 4, 7, 8, 9]                                   # equal scores: lower row id first
    assert i[1].tolist() == [6, 5, 0, 3, 8, 9]
    s, i = ops.topk_ip(Q.to(cuda_dev), P[:4].to(cuda_dev), 6)                    # fewer passages than K: -1 padding
    assert i[0].tolist() == [1, 2, 3, 0, -1, -1] and torch.isinf(s[0, 4:]).all()
    big = torch.randn(50, 64).to(cuda_dev)
    view = big[:, :32]                                                           # row stride 64, 32 used columns
    s1, i1 = ops.topk_ip(Q[:, :0].new_zeros(1, 32).to(cuda_dev) + 1.0, view, 5)
    s2, i2 = ops.topk_ip(torch.ones(1, 32, device=cuda_dev), view.contiguous(), 5)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)


def test_exact_index_speaks_hnswlib(cuda_dev):
    from dalm_b200.eval.utils import construct_search_index, get_nearest_neighbours
    from oracle import topk
    g = np.random.default_rng(3)
    data = g.standard_normal((500, 128)); data /= np.linalg.norm(data, axis=1, keepdims=True)
    q = g.standard_normal((7, 128)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    index = construct_search_index(128, 500, data)
    labels, dist = index.knn_query(q, k=10)
    want_l, want_d = topk.knn_query(data.astype(np.float32), q.astype(np.float32), 10)
    assert np.array_equal(labels, want_l) and np.allclose(dist, want_d, atol=2e-6)
    ids = {i: f"passage {i}" for i in range(500)}
    res = get_nearest_neighbours(10, index, q, ids, threshold=0.0)
    for row, wl, wd in zip(res, want_l, want_d):
        keep = [(f"passage {l}", 1 - d) for l, d in zip(wl, wd) if 1 - d >= 0.0]
        assert [p for p, _ in row] == [p for p, _ in keep]
    with pytest.raises(RuntimeError):
        index.knn_query(q, k=501)                                                # hnswlib raises when k > element count
    with pytest.raises(RuntimeError):
        index.add_items(data[:1])


def test_retriever_metrics_on_batch_match_oracle_pipeline(cuda_dev):
    """same embeddings -> identical precision / recall / hit lists as the reference's per-batch routine over an exact index"""
    from dalm_b200.eval.utils import calc_eval_results, construct_search_index, evaluate_retriever_on_batch
    from oracle import topk
    g = torch.Generator().manual_seed(9)
    n_pass, D, K = 60, 64, 5
    table = torch.nn.functional.normalize(torch.randn(200, D, generator=g), dim=1)           # token id -> embedding
    fwd = lambda ids, mask: table.to(ids.device)[ids[:, 0]]
    passages = [f"p{i}" for i in range(n_pass)]
    pemb = table[:n_pass].numpy()
    index = construct_search_index(D, n_pass, pemb)
    id_to_passage = dict(enumerate(passages))
    # queries: noisy copies of their passage's embedding row ids (some wrong on purpose)
    q_tok = torch.tensor([[i if i % 4 else (i + 100)] for i in range(n_pass)])
    batch = {"retriever_query_input_ids": q_tok, "retriever_query_attention_mask": torch.ones_like(q_tok), "Abstract": passages}
    prec, rec, hits, top = evaluate_retriever_on_batch(batch, "Abstract", fwd, index, torch.bfloat16, str(cuda_dev), K, id_to_passage)
    labels, dist = topk.knn_query(pemb, table[q_tok[:, 0]].numpy(), K)
    w_prec, w_rec, w_hits, w_top = [], [], 0, []
    for i in range(n_pass):
        got = [passages[l] for l, d in zip(labels[i], dist[i]) if 1 - d >= 0.0]
        w_top.append(got[0])
        c = len(set(got) & {passages[i]})
        w_prec.append(c / len(set(got))); w_rec.append(float(c)); w_hits += passages[i] in got
    assert prec == w_prec and rec == w_rec and hits == w_hits and top == w_top
    r = calc_eval_results(n_pass, prec, rec, hits)
    assert 0.7 < r.recall < 0.8 and r.hit_rate == r.recall


def test_evaluate_retriever_end_to_end_with_trained_adapters(cuda_dev, tmp_path):
    """`dalm train-retriever-only` then `dalm eval-retriever --retriever-peft-model-path ...` on the toy CSV: the evaluation
    wrapper (built frozen, adapters attached afterwards like PeftModel.from_pretrained) embeds exactly like the trained model"""
    from dalm_b200 import synthetic
    from dalm_b200.eval.eval_rag import evaluate_rag
    from dalm_b200.eval.eval_retriever_only import evaluate_retriever
    from dalm_b200.models.rag_e2e_base_model import inference_only
    from dalm_b200.models.retriever_only_base_model import AutoModelForSentenceEmbedding
    from dalm_b200.training.retriever_only.train_retriever_only import train_retriever
    csv = synthetic.write_csv(str(tmp_path / "toy.csv"), 24, seed=5)
    rdir = synthetic.write_model_dir(str(tmp_path / "bge-tiny"), "bert", "bge-tiny", vocab_size=1200)
    out = str(tmp_path / "out")
    res = train_retriever(rdir, csv, per_device_train_batch_size=4, query_max_len=16, passage_max_len=32, num_train_epochs=2,
                          output_dir=out, use_peft=True, use_bnb=False, with_tracking=False, learning_rate=2e-3)
    base = evaluate_retriever(csv, rdir, None, "Abstract", "Question", embed_dim=64, max_length=32, test_batch_size=8, top_k=5)
    tuned = evaluate_retriever(csv, rdir, os.path.join(out, "retriever"), "Abstract", "Question", embed_dim=64, max_length=32,
                               test_batch_size=8, top_k=5)
    for r in (base, tuned):
        assert r.total_examples == 24 and 0.0 <= r.recall <= 1.0 and 0.0 <= r.precision <= r.recall and r.hit_rate == r.recall
    assert tuned.recall >= base.recall                                              # two epochs of contrastive training on this set
    # adapters attached to a frozen-built wrapper == a wrapper built with get_peft and the same adapter weights
    ids = torch.randint(5, 1200, (3, 32)); mask = torch.ones_like(ids)
    with inference_only():
        m1 = AutoModelForSentenceEmbedding(rdir, get_peft=False, use_bnb=False)
    assert m1.model.full is None and m1.model.lora is None
    m1.attach_pre_trained_peft_layers(os.path.join(out, "retriever"), "cuda")
    m2 = AutoModelForSentenceEmbedding(rdir, get_peft=True, use_bnb=False)
    m2.attach_pre_trained_peft_layers(os.path.join(out, "retriever"), "cuda")
    with torch.no_grad():
        assert torch.equal(m1(ids, mask), m2(ids, mask))
    with pytest.raises(NotImplementedError):
        evaluate_rag(csv, rdir, rdir, None, None, "Abstract", "Question", "Answer", 64, 32)      # a BERT encoder is no generator


def test_evaluate_rag_with_generator(cuda_dev, tmp_path, capsys):
    """`dalm eval-rag` with generator evaluation on (the reference's default): retrieve, build `#query# .. #passage# .. #answer# `
    prompts from the top passage, decode greedily through the KV-cache kernels, score exact match (reference
    eval_rag.py:126-164,258-283). Token-level parity of the decoding is in tests/test_generate_gpu.py; here the plumbing:
    prompt text survives, the continuation is appended after it, batches of `query_batch_size` plus the leftover are all
    generated, and a prompt as long as max_length is rejected exactly like HF rejects it."""
    import csv as _csv

    from dalm_b200 import synthetic
    from dalm_b200.eval.eval_rag import evaluate_rag, run_generator_on_prompts
    from dalm_b200.models.rag_e2e_base_model import AutoModelForRagE2E, inference_only
    words = synthetic.word_list()
    path = str(tmp_path / "short.csv")
    with open(path, "w", newline="") as f:
        w = _csv.DictWriter(f, fieldnames=["Abstract", "Question", "Answer"])
        w.writeheader()
        for i in range(11):                                                        # 11 rows: batches of 4 + a leftover
            w.writerow({"Abstract": " ".join(words[20 + 6 * i:26 + 6 * i]), "Question": " ".join(words[200 + 4 * i:204 + 4 * i]),
                        "Answer": " ".join(words[400 + i:402 + i])})
    rdir = synthetic.write_model_dir(str(tmp_path / "bge-tiny"), "bert", "bge-tiny", vocab_size=1200)
    gdir = synthetic.write_model_dir(str(tmp_path / "llama-tiny"), "llama", "llama-tiny", vocab_size=1200)
    res = evaluate_rag(path, rdir, gdir, None, None, "Abstract", "Question", "Answer", embed_dim=64, max_length=96,
                       test_batch_size=4, query_batch_size=4, top_k=3, evaluate_generator=True)
    out = capsys.readouterr().out
    assert res.total_examples == 11 and 0.0 <= res.recall <= 1.0
    assert "Generator evaluation:" in out and "Exact match: 0.0" in out            # a random-init generator answers nothing
    with inference_only():
        rag = AutoModelForRagE2E(rdir, gdir)
    tok = rag.generator_tokenizer
    tok.pad_token = tok.eos_token
    prompts = [f"#query# {' '.join(words[i:i + 4])} #passage# {' '.join(words[30 + i:36 + i])} #answer# " for i in range(3)]
    texts = run_generator_on_prompts(rag.generator_model, tok, prompts, max_length=96)
    assert len(texts) == 3
    for p, t in zip(prompts, texts):
        shown = tok.decode(tok(p)["input_ids"], skip_special_tokens=True)          # the prompt as the tokenizer round-trips it
        assert t.startswith(shown.strip()) and len(t) > len(shown.strip())         # prompt kept, something generated after it
    with pytest.raises(ValueError):
        run_generator_on_prompts(rag.generator_model, tok, prompts, max_length=8)   # truncated prompt fills max_length (HF raises too)
