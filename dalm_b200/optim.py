"""Fused Adam over the flat parameter buffers (LoRA banks, or the fp32 master buffer of a fully fine-tuned model together
with its bf16 shadow) — torch.optim.Adam semantics (reference train_rage2e.py:336: lr, betas (0.9, 0.999), eps 1e-8, no
weight decay), one kernel launch per buffer."""
from __future__ import annotations

from typing import Iterable

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                bank = getattr(p, "_dalm_bank", None)
                if bank is not None:                  # full fine-tuning: update master + refresh the bf16 shadow in one pass
                    ops.adam_step_shadow_(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], bank.p16, group["lr"], b1, b2,
                                          group["eps"], st["step"])
                elif p.is_cuda and p.dtype == torch.float32 and p.is_contiguous():
                    ops.adam_step_(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"], st["step"])
                else:
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
        return None

    def zero_grad(self, set_to_none: bool = False) -> None:
        # gradients live in persistent flat buffers that the kernels accumulate into: zero in place, never drop them
        for group in self.param_groups:
            for p in group["params"]:
                bank = getattr(p, "_dalm_bank", None)
                if bank is not None:
                    bank.zero_grad()                  # weight gradients are overwritten by the next wgrad: no 4 B/param memset
                elif p.grad is not None:
                    p.grad.zero_()
