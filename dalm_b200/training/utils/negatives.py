"""Cross-rank in-batch negatives (optional extension, OFF by default).

The reference keeps in-batch negatives rank-local: `get_cosine_sim(query_embs, passage_embs)` sees only this rank's B rows
(train_rage2e.py:431-446, train_retriever_only.py:367-373), so an 8-GPU run contrasts each query against 18 passages, not 144.
With DALM_B200_CROSS_RANK_NEGATIVES=1 every rank all-gathers the pooled embeddings (W x B x D fp32 - 590 KB at cfg-4, one small
NCCL collective over NVLink) and runs the SAME fused in-batch kernel on the [W B, D] matrices:

    J = (1/W) sum_r L_r ,   L_r = Lc(S_global) + doc_r + tok_r ,   S_global = s Q_all P_all^T  [W B, W B]
    doc_r = - sum_{b on rank r} c_b log_softmax(S_global[b,:])[b] / N_r

Lc is the two-way contrastive loss over all W B rows (identical on every rank); the marginalisation's doc term of a row uses
the global candidate set. Each rank keeps rows [rank B, rank B + B) of dQ_all / dP_all and back-propagates them through ITS
encoder pass; they are scaled by W because the data-parallel gradient exchange then takes the MEAN over ranks, which turns
sum_r (W dJ/dq_r)(dq_r/dtheta) into exactly dJ/dtheta. No gradient flows "through" the all-gather: every rank computes the
full dS itself (B^2 W^2 exps: 21 k at 8 x 18), which is cheaper than a reduce-scatter of dP_all.

The step then runs eagerly (a NCCL collective in the middle of the launch sequence is not captured into the step's CUDA graph).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch

f32 = torch.float32


def enabled() -> bool:
    return os.environ.get("DALM_B200_CROSS_RANK_NEGATIVES", "0") == "1"


def active() -> bool:
    """cross-rank negatives requested AND more than one rank in the job"""
    import torch.distributed as dist
    return enabled() and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def dist_row_counts(n_rows: int, device) -> List[int]:
    """every rank's local batch height (a short final batch on one rank makes them differ): one tiny all-gather + host read"""
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([n_rows], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    return [int(c) for c in torch.cat(counts).tolist()]


def dist_gather_rows(t: torch.Tensor, counts: Optional[List[int]] = None) -> List[torch.Tensor]:
    """all-gather of per-rank row blocks whose heights may differ: -> list of W tensors (rank order). `counts`: the heights, if
    the caller already exchanged them (one exchange serves all gathers of a step)"""
    import torch.distributed as dist
    world = dist.get_world_size()
    if counts is None:
        counts = dist_row_counts(t.shape[0], t.device)
    mx = max(counts)
    pad = t if t.shape[0] == mx else torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return [b[:c] for b, c in zip(bufs, counts)]


def global_inbatch_loss(q_emb: torch.Tensor, p_emb: torch.Tensor, logit_scale: float, cvec: Optional[torch.Tensor],
                        nsum: Optional[torch.Tensor], need_grad: bool, grad_out: float, *, rank: int, world: int,
                        loss_fn: Callable, gather: Optional[Callable[[torch.Tensor], List[torch.Tensor]]] = None) -> Dict[str, torch.Tensor]:
    """Same contract as ops.inbatch_loss for the LOCAL rows (S is the global matrix): dict(S, dlp, losses, dQ, dP).
    loss_fn: ops.inbatch_loss (injectable so the host logic is testable on CPU); gather: rows of every rank, rank-major
    (default: torch.distributed all-gathers sharing one exchange of the batch heights)."""
    if gather is None:
        counts = dist_row_counts(q_emb.shape[0], q_emb.device)
        gather = lambda t: dist_gather_rows(t, counts)
    qs, ps = gather(q_emb.contiguous()), gather(p_emb.contiguous())
    q_all, p_all = torch.cat(qs, 0), torch.cat(ps, 0)
    lo = sum(t.shape[0] for t in qs[:rank])
    B = q_emb.shape[0]
    cvec_all = None
    if cvec is not None:
        # per-row weight c_b / N_rank(b); the kernel divides by `nsum` (this rank's N) and the (1/W) of J's mean is folded in
        w_all = torch.cat(gather((cvec / nsum).contiguous()), 0)
        cvec_all = (w_all * (nsum / float(world))).contiguous()
    r = loss_fn(q_all, p_all, float(logit_scale), cvec_all, nsum, need_grad=need_grad, grad_out=float(grad_out) * world)
    dlp = r["dlp"][lo:lo + B]
    lc = r["losses"][0]
    if cvec is not None:
        doc = -(cvec * dlp).sum() / nsum[0]
        n = nsum[0]
    else:
        doc = torch.zeros((), dtype=f32, device=q_emb.device)
        n = torch.zeros((), dtype=f32, device=q_emb.device)
    losses = torch.stack([lc, doc, lc + doc, n]).to(f32)
    return {"S": r["S"], "dlp": dlp, "losses": losses,
            "dQ": r["dQ"][lo:lo + B].contiguous() if need_grad else None,
            "dP": r["dP"][lo:lo + B].contiguous() if need_grad else None}
