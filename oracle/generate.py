"""CPU restatement of HF greedy `generate` as the reference calls it (dalm/eval/eval_rag.py:126-140:
`model.generate(**inputs, max_length=max_length, early_stopping=True)`). TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

transformers is a third-party dependency of the reference (pyproject.toml: `transformers>4.35`, unpinned); the installed
5.5.0 is the pin here: tests/test_generate_host.py checks this restatement token for token against
`LlamaForCausalLM.generate` / `FalconForCausalLM.generate` of that version (left- and right-padded prompts, EOS reached and
not reached). What it states (GenerationMixin._sample, do_sample=False):
  * position ids = cumsum(attention_mask) - 1 (1 where the mask is 0), recomputed as ones are appended to the mask;
  * next token = argmax of the last column's logits; rows that already emitted EOS emit pad_token_id (= first EOS id when
    the model has no pad token) from then on;
  * the loop ends right after the step in which the last row emitted EOS, or at max_length TOTAL tokens (prompt included).
No KV cache: every step re-runs the whole prefix, which is the definition the cache must reproduce.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


@torch.no_grad()
def greedy_generate(model, input_ids: torch.Tensor, attention_mask: torch.Tensor, max_length: int,
                    eos_token_ids: Sequence[int] = (), pad_token_id: Optional[int] = None) -> torch.Tensor:
    if input_ids.shape[1] >= max_length:
        raise ValueError(f"Input length of input_ids is {input_ids.shape[1]}, but `max_length` is set to {max_length}.")
    eos = [int(e) for e in eos_token_ids]
    pad = int(pad_token_id) if pad_token_id is not None else (eos[0] if eos else 0)
    toks, am = input_ids.clone().long(), attention_mask.clone().long()
    unfinished = torch.ones(toks.shape[0], dtype=torch.long)
    while toks.shape[1] < max_length:
        pos = (am.cumsum(-1) - 1).masked_fill(am == 0, 1)
        logits = model(input_ids=toks, attention_mask=am, position_ids=pos).logits[:, -1].float()
        nxt = logits.argmax(-1)
        nxt = nxt * unfinished + pad * (1 - unfinished)
        toks = torch.cat([toks, nxt[:, None]], 1)
        am = torch.cat([am, torch.ones(am.shape[0], 1, dtype=torch.long)], 1)
        for e in eos:
            unfinished = unfinished & (nxt != e).long()
        if eos and int(unfinished.max()) == 0:
            break
    return toks


@torch.no_grad()
def step_margins(model, tokens: torch.Tensor, attention_mask: torch.Tensor, prompt_len: int) -> List[torch.Tensor]:
    """teacher-forced check of a generated continuation: for every generated column c >= prompt_len returns, per row,
    (best logit - logit of the token actually emitted at c) under THIS model given the emitted prefix. 0 = the emitted token
    is the oracle's argmax; a bf16 implementation may legitimately pick a token whose oracle margin is within its rounding
    noise. (Rows already finished emit pad and are not meaningful; the caller masks them.)"""
    B, T = tokens.shape
    am = torch.cat([attention_mask.long(), torch.ones(B, T - attention_mask.shape[1], dtype=torch.long)], 1)
    pos = (am.cumsum(-1) - 1).masked_fill(am == 0, 1)
    logits = model(input_ids=tokens, attention_mask=am, position_ids=pos).logits.float()
    out = []
    for c in range(prompt_len, T):
        row = logits[:, c - 1]
        out.append(row.max(-1).values - row.gather(1, tokens[:, c:c + 1]).squeeze(1))
    return out
