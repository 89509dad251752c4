"""LoRA adapters (PEFT semantics: y = W x + (alpha/r) * B(A(x)); r=8, alpha=16, dropout 0.05, bias none — reference
dalm/models/rag_e2e_base_model.py:144-160) stored as ONE flat fp32 buffer per model so that the optimizer and the
data-parallel all-reduce each touch a single contiguous range."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch


class LoraBank:
    def __init__(self, specs: List[Tuple[str, int, int]], r: int = 8, alpha: int = 16, dropout: float = 0.05,
                 device="cuda", seed: int = 0):
        """specs: (module_name, in_features, out_features) for every adapted Linear, in a fixed order."""
        self.r, self.alpha, self.dropout = r, alpha, dropout
        self.scale = alpha / r
        self.specs = specs
        self.param = None                 # the nn.Parameter over `flat` (set by the engine model): its .grad follows rebind_grad
        total = sum(r * i + o * r for _, i, o in specs)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.A: Dict[str, torch.Tensor] = {}
        self.B: Dict[str, torch.Tensor] = {}
        self.gA: Dict[str, torch.Tensor] = {}
        self.gB: Dict[str, torch.Tensor] = {}
        off = 0
        gen = torch.Generator().manual_seed(seed)
        for name, fin, fout in specs:
            self.A[name] = self.flat[off:off + r * fin].view(r, fin)
            self.gA[name] = self.grad[off:off + r * fin].view(r, fin)
            bound = 1.0 / math.sqrt(fin)          # kaiming_uniform_(a=sqrt(5)) on [r, fin]  (PEFT default init)
            init = (torch.rand(r, fin, generator=gen) * 2 - 1) * bound
            self.A[name].copy_(init)
            off += r * fin
            self.B[name] = self.flat[off:off + fout * r].view(fout, r)      # zeros (PEFT default)
            self.gB[name] = self.grad[off:off + fout * r].view(fout, r)
            off += fout * r

    def numel(self) -> int:
        return self.flat.numel()

    def rebind_grad(self, buf: torch.Tensor) -> None:
        """move the flat gradient buffer (and its per-adapter views) into `buf` — a slice of the data-parallel gradient
        arena, so that all trainable banks + the loss scalar are ONE contiguous all-reduce (accel.GradientSync). Must happen
        before a step is captured into a CUDA graph (the kernels' output pointers change)."""
        if buf.numel() != self.grad.numel() or buf.dtype != self.grad.dtype or not buf.is_contiguous():
            raise ValueError("rebind_grad: buffer must be a contiguous fp32 tensor of the bank's size")
        buf.copy_(self.grad)
        self.grad = buf
        off, r = 0, self.r
        for name, fin, fout in self.specs:
            self.gA[name] = buf[off:off + r * fin].view(r, fin)
            off += r * fin
            self.gB[name] = buf[off:off + fout * r].view(fout, r)
            off += fout * r
        if self.param is not None:
            self.param.grad = buf

    def zero_grad(self) -> None:
        self.grad.zero_()

    # PEFT adapter state-dict naming: base_model.model.<module>.lora_A.weight / lora_B.weight
    def peft_state_dict(self, prefix: str = "base_model.model.") -> Dict[str, torch.Tensor]:
        out = {}
        for name, _, _ in self.specs:
            out[f"{prefix}{name}.lora_A.weight"] = self.A[name].detach().clone().cpu()
            out[f"{prefix}{name}.lora_B.weight"] = self.B[name].detach().clone().cpu()
        return out

    def load_peft_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "base_model.model.") -> None:
        for name, _, _ in self.specs:
            for which, store in (("lora_A", self.A), ("lora_B", self.B)):
                for key in (f"{prefix}{name}.{which}.weight", f"{prefix}{name}.{which}.default.weight"):
                    if key in sd:
                        store[name].copy_(sd[key].to(store[name].device, torch.float32))
                        break
                else:
                    raise KeyError(f"adapter weight for {name}.{which} not found")
