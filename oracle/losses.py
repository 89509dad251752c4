"""CPU restatement of the loss functions of reference dalm/training/utils/train_utils.py:76-138.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Two forms are given for the marginalised loss: `*_loopform`
follows the reference line by line (slice / add / cat / stack / gather), `*_closed` is the closed form derived in
SURVEY.md §8a that the CUDA kernels implement. tests/test_oracle_golden.py checks both against the committed outputs of
the reference itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def get_cosine_sim(q: torch.Tensor, p: torch.Tensor, logit_scale: float) -> torch.Tensor:
    """train_utils.py:76-77"""
    return (q @ p.t()) * logit_scale


def get_nt_xent_loss(sim: torch.Tensor) -> torch.Tensor:
    """train_utils.py:80-88: CE of each row against its own index, mean-reduced"""
    n = sim.shape[0]
    lsm = sim - torch.logsumexp(sim, dim=1, keepdim=True)
    return -lsm[torch.arange(n), torch.arange(n)].mean()


def contrastive_loss(sim: torch.Tensor) -> torch.Tensor:
    """loop body train_rage2e.py:443-446 / train_retriever_only.py:371-373"""
    return (get_nt_xent_loss(sim) + get_nt_xent_loss(sim.t())) / 2.0


def marginalized_loss_loopform(logits, input_ids, attention_mask, scores, qlen) -> torch.Tensor:
    """train_utils.py:113-138, statement by statement"""
    B, L, V = logits.shape
    lp = F.log_softmax(logits[:, :-1, :], dim=2).view(B, -1, V)                       # :121
    doc = torch.log_softmax(scores, dim=1).diag().unsqueeze(-1).unsqueeze(-1)          # :124
    out = []
    for i in range(B):                                                                 # :127-131
        q = int(qlen[i])
        sl = lp[i]
        head = sl[: q - 1, :]                                                          # :101
        tail = sl[q - 1:, :] + doc[i]                                                  # :104-106
        out.append(torch.cat([head, tail], dim=0))                                     # :109
    m = torch.stack(out)                                                               # :133
    nll = -torch.gather(m, 2, input_ids[:, 1:].unsqueeze(2)).squeeze(-1)               # :91-93,134
    w = attention_mask[:, 1:]
    return (nll * w).sum() / w.sum()                                                   # :135-136


def marginal_counts(attention_mask: torch.Tensor, qlen: torch.Tensor):
    """c_b = sum_t m[b,t+1]*[t >= start_b], N = sum m[:,1:], with python slice semantics for start_b = qlen_b-1"""
    B, L = attention_mask.shape
    m = attention_mask[:, 1:].to(torch.float64)
    t = torch.arange(L - 1).unsqueeze(0)
    start = qlen.to(torch.int64) - 1
    start = torch.where(start < 0, torch.clamp(start + (L - 1), min=0), start)
    ind = (t >= start.unsqueeze(1)).to(torch.float64)
    return (m * ind).sum(1), m.sum()


def marginalized_loss_closed(logits, input_ids, attention_mask, scores, qlen):
    """SURVEY §8a closed form. returns (Lm, tok_lp [B,L-1], dlp [B], c [B], N)"""
    lp = F.log_softmax(logits[:, :-1, :].to(torch.float64), dim=2)
    tok = torch.gather(lp, 2, input_ids[:, 1:].unsqueeze(2)).squeeze(-1)
    dlp = torch.log_softmax(scores.to(torch.float64), dim=1).diag()
    c, N = marginal_counts(attention_mask, qlen)
    m = attention_mask[:, 1:].to(torch.float64)
    Lm = -((m * tok).sum() + (c * dlp).sum()) / N
    return Lm, tok, dlp, c, N


def rag_loss_and_grads(q, p, logit_scale, logits, input_ids, attention_mask, qlen):
    """Full combined loss of train_rage2e.py:441-467 with autograd gradients (fp64 on CPU).
    returns dict(loss, Lc, Lm, S, dQ, dP, dlogits)"""
    q = q.detach().to(torch.float64).requires_grad_(True)
    p = p.detach().to(torch.float64).requires_grad_(True)
    lg = logits.detach().to(torch.float64).requires_grad_(True)
    S = get_cosine_sim(q, p, logit_scale)
    Lc = contrastive_loss(S)
    Lm = marginalized_loss_loopform(lg, input_ids, attention_mask, S, qlen)
    loss = Lc + Lm
    loss.backward()
    return {"loss": loss.detach(), "Lc": Lc.detach(), "Lm": Lm.detach(), "S": S.detach(), "dQ": q.grad, "dP": p.grad,
            "dlogits": lg.grad}


def retriever_loss_and_grads(q, p, logit_scale):
    """train_retriever_only.py:365-376"""
    q = q.detach().to(torch.float64).requires_grad_(True)
    p = p.detach().to(torch.float64).requires_grad_(True)
    S = get_cosine_sim(q, p, logit_scale)
    loss = contrastive_loss(S)
    loss.backward()
    return {"loss": loss.detach(), "S": S.detach(), "dQ": q.grad, "dP": p.grad}
