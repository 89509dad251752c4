"""torch.autograd bridges: make the engine's hand-written forward/backward visible to callers that drive the step the
way the reference loop does (`loss.backward()`), without any arithmetic in PyTorch."""
from __future__ import annotations

import torch

from .. import ops


class EncodeFn(torch.autograd.Function):
    """ids, mask -> pooled (+normalised) sentence embeddings [B,H] fp32.  `anchor` is the encoder's flat LoRA parameter:
    it ties the output to the autograd graph; its gradient is accumulated in place by the kernels (anchor.grad is the
    LoRA bank's gradient buffer), so backward returns None for it."""

    @staticmethod
    def forward(ctx, anchor, enc, ids, mask, normalize, pool_mask=None):
        hid, c = enc.forward_hidden(ids, mask, save=True)
        pm = mask if pool_mask is None else pool_mask          # autoregressive retrievers pool with the eos one-hot mask
        emb, norm = ops.pool_norm_fwd(hid, pm, normalize)
        ctx.enc, ctx.c, ctx.mask, ctx.normalize, ctx.L = enc, c, pm, normalize, ids.shape[1]
        ctx.save_for_backward(emb, norm)
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        emb, norm = ctx.saved_tensors
        d_hid = ops.pool_norm_bwd(emb, norm, d_emb.contiguous().float(), ctx.mask, ctx.L, ctx.normalize)
        ctx.enc.backward_hidden(ctx.c, d_hid)
        ctx.c = None
        return None, None, None, None, None, None


class GenerateFn(torch.autograd.Function):
    """ids, mask -> logits [B,L,V] bf16."""

    @staticmethod
    def forward(ctx, anchor, dec, ids, mask):
        logits, c = dec.forward_logits(ids, mask, save=True)
        ctx.dec, ctx.c = dec, c
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.dec.backward_logits(ctx.c, dlogits.contiguous().to(torch.bfloat16))
        ctx.c = None
        return None, None, None, None


class PoolFn(torch.autograd.Function):
    """mean_pooling (no normalisation) as a stand-alone differentiable op for API users"""

    @staticmethod
    def forward(ctx, hidden, mask):
        emb, norm = ops.pool_norm_fwd(hidden.float().contiguous(), mask, False)
        ctx.mask, ctx.L = mask, hidden.shape[1]
        ctx.save_for_backward(emb, norm)
        return emb

    @staticmethod
    def backward(ctx, d):
        emb, norm = ctx.saved_tensors
        return ops.pool_norm_bwd(emb, norm, d.contiguous().float(), ctx.mask, ctx.L, False), None
