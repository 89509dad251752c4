"""CPU: oracle/generate.py (the restatement the GPU decode path is checked against) is pinned token for token to the installed
transformers' own `generate` — the call the reference makes (dalm/eval/eval_rag.py:136-139) — for Llama and Falcon, with
left- and right-padded prompts, with EOS reached (rows finish at different steps and are padded) and never reached."""
import warnings

import pytest
import torch

from dalm_b200 import synthetic
from dalm_b200.engine import params
from oracle import generate as og
from oracle import models as om


def _prompt(B, L0, V, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, V, (B, L0), generator=g)
    mask = torch.ones(B, L0, dtype=torch.int64)
    mask[1, :3] = 0          # left padding
    mask[2, L0 - 3:] = 0     # right padding (HF warns, the reference's tokenizer setting decides which one it gets)
    return ids, mask


def _hf_generate(model, ids, mask, T, eos, pad):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return model.generate(input_ids=ids, attention_mask=mask, max_length=T, early_stopping=True, do_sample=False,
                              eos_token_id=eos, pad_token_id=pad)


@pytest.mark.parametrize("family,name", [("llama", "llama-tiny"), ("falcon", "falcon-tiny")])
def test_restatement_matches_hf_generate(family, name):
    V = 400
    cfg = synthetic.llama_config(name, V) if family == "llama" else synthetic.falcon_config(name, V)
    model = (om.build_llama if family == "llama" else om.build_falcon)(cfg, params.random_state_dict(family, cfg, seed=2))
    ids, mask = _prompt(4, 10, V, seed=1)
    T = 26
    free = _hf_generate(model, ids, mask, T, None, 0)                       # no EOS: runs to max_length
    assert free.shape == (4, T)
    got = og.greedy_generate(model, ids, mask, T, eos_token_ids=(), pad_token_id=0)
    assert torch.equal(got, free)
    # choose EOS ids among tokens that the free run emits at different steps so that rows finish at different times
    eos = [int(free[0, 13]), int(free[3, 17])]
    want = _hf_generate(model, ids, mask, T, eos, eos[0])
    got = og.greedy_generate(model, ids, mask, T, eos_token_ids=eos, pad_token_id=eos[0])
    assert want.shape == got.shape and torch.equal(got, want)
    # a single early EOS that every row hits must stop generation before max_length, or rows keep padding until then
    margins = og.step_margins(model, free, mask, 10)
    assert len(margins) == T - 10 and all(float(m.abs().max()) == 0.0 for m in margins)   # the free run is its own argmax


def test_prompt_at_max_length_is_rejected():
    cfg = synthetic.llama_config("llama-tiny", 400)
    model = om.build_llama(cfg, params.random_state_dict("llama", cfg, seed=2))
    ids, mask = _prompt(4, 10, 400, seed=1)
    with pytest.raises(ValueError):
        og.greedy_generate(model, ids, mask, 10)
    with pytest.raises(ValueError):
        _hf_generate(model, ids, mask, 10, None, 0)


# ----------------------------------------------------------------------------------------------------------------
# host logic of dalm_b200/engine/decoding.py with the three device kernels emulated in torch (their C-ABI contract,
# include/dalm_b200.h) and a decoder stub that scores with the HF oracle: column bookkeeping, position ids, the
# every-8-tokens alive check, trimming after the last EOS and the HF argument conventions run on CPU here; the kernels
# themselves are covered by tests/test_generate_gpu.py
# ----------------------------------------------------------------------------------------------------------------
class _StubDecoder(torch.nn.Module):
    def __init__(self, hf_model, cfg):
        super().__init__()
        self.m, self.cfg, self.V, self.dev = hf_model, cfg, cfg["vocab_size"], torch.device("cpu")
        self.layers, self.lm_head = [None, None], None
        self.seen_pos = []

    def _rope(self, total):
        return (torch.zeros(total, 1), torch.zeros(total, 1))

    def kv_columns(self):
        return 0, 8, 8

    def _scores(self):
        pos = (self.am.cumsum(-1) - 1).masked_fill(self.am == 0, 1)
        with torch.no_grad():
            return self.m(input_ids=self.toks, attention_mask=self.am, position_ids=pos).logits[:, -1].to(torch.bfloat16), pos

    def _prefill_last(self, ids, mask, pos, tables, sink):
        self.toks, self.am = ids.clone(), mask.clone()
        logits, want = self._scores()
        assert torch.equal(pos.view_as(want), want)                          # prompt position ids = cumsum(mask) - 1
        for li in range(len(self.layers)):
            sink(li, torch.zeros(ids.numel(), 16, dtype=torch.bfloat16))     # exercises the cache copy shapes
        return logits

    def _decode_step(self, ids, pos, caches, kmask, cur, tables):
        assert cur.dtype == torch.int32 and (cur == self.toks.shape[1]).all()    # K / V of this token go to the next free column
        cur = int(cur[0])
        self.toks = torch.cat([self.toks, ids[:, None]], 1)
        self.am = torch.cat([self.am, torch.ones(len(ids), 1, dtype=torch.int64)], 1)
        assert torch.equal(kmask[:, :cur + 1], self.am)                      # the device mask has grown with the tokens
        logits, want = self._scores()
        assert torch.equal(pos, want[:, -1])                                 # position id of the token being decoded
        return logits


def _emulated_greedy_step(logits, V, eos_ids, pad_id, unfinished, tokens, mask, col, next_ids, pos, alive):
    """torch restatement of dalm_b200_greedy_step's contract (include/dalm_b200.h), row by row like the kernel's CTAs"""
    cur_dev = col if torch.is_tensor(col) else None
    T = tokens.shape[1]
    eos = [] if eos_ids is None else eos_ids.tolist()
    for b in range(logits.shape[0]):
        c = int(cur_dev[b]) + 1 if cur_dev is not None else int(col)
        if c >= T:
            continue                                                         # a replay past the end of the buffers is a no-op
        tok = int(logits[b, :V].float().argmax()) if int(unfinished[b]) else int(pad_id)
        tokens[b, c], mask[b, c], next_ids[b] = tok, 1, tok
        pos[b] += 1
        if int(unfinished[b]) and tok in eos:
            unfinished[b] = 0
        alive[c] += int(unfinished[b])
        if cur_dev is not None:
            cur_dev[b] = c


@pytest.mark.parametrize("use_eos", [False, True])
def test_host_loop_with_emulated_kernels(monkeypatch, use_eos):
    from dalm_b200 import ops
    from dalm_b200.engine.decoding import greedy_generate

    cfg = synthetic.llama_config("llama-tiny", 400)
    model = om.build_llama(cfg, params.random_state_dict("llama", cfg, seed=2))
    ids, mask = _prompt(4, 10, 400, seed=1)
    T = 31                                                                   # > 2 x 8 generated tokens: several alive checks
    free = og.greedy_generate(model, ids, mask, T, eos_token_ids=(), pad_token_id=0)
    eos = [int(free[0, 12]), int(free[3, 14]), int(free[1, 11]), int(free[2, 13])] if use_eos else None
    want = og.greedy_generate(model, ids, mask, T, eos_token_ids=eos or (), pad_token_id=(eos[0] if eos else 7))
    if use_eos:
        assert want.shape[1] < T                                             # every row finishes early: the trim path runs
    monkeypatch.setattr(ops, "gemm_rows", lambda a, b, **k: a)              # the stub decoder already returns logits
    monkeypatch.setattr(ops, "greedy_step_", _emulated_greedy_step)
    dec = _StubDecoder(model, dict(cfg, eos_token_id=None))
    dec.train()
    got = greedy_generate(dec, input_ids=ids, attention_mask=mask, max_length=T, early_stopping=True,
                          eos_token_id=eos, pad_token_id=None if use_eos else 7, token_type_ids=None)
    assert dec.training                                                      # mode restored
    assert got.shape == want.shape and torch.equal(got, want)
    with pytest.raises(ValueError):
        greedy_generate(dec, input_ids=ids, attention_mask=mask, max_length=10)
    with pytest.raises(NotImplementedError):
        greedy_generate(dec, input_ids=ids, attention_mask=mask, max_length=T, do_sample=True)
    with pytest.raises(NotImplementedError):
        greedy_generate(dec, input_ids=ids, attention_mask=mask, max_length=8193)       # beyond the decode kernel's cache limit
    short = greedy_generate(dec, input_ids=ids, attention_mask=mask, max_new_tokens=3, eos_token_id=None, pad_token_id=7)
    assert torch.equal(short, free[:, :13])
