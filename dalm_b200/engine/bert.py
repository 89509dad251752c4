"""BERT encoder (bge-small / bge-large shapes) forward + backward as a launch sequence over the C-ABI kernels.

Mirrors what `self.retriever_model(input_ids, attention_mask)[0]` computes in the reference
(dalm/models/rag_e2e_base_model.py:93, dalm/models/retriever_only_base_model.py:58) through HF BertModel:
embeddings(word+pos+type0) -> LN -> N x [QKV(+LoRA) -> masked softmax attention -> out-proj + residual -> LN ->
FFN(GELU erf) + residual -> LN], eps 1e-12, token_type_ids = 0 (the reference calls the model positionally).

HBM layout (per layer, bf16 unless noted):
  Wqkv_aug [3H, H+Ra]   rows q|k|v of the fused projection; the last Ra=3r columns hold (alpha/r)*B_j so that LoRA's
                        up-projection is part of the same tcgen05 GEMM (A operand = [x | x A^T])
  WqkvT_aug [H, 3H+Ra]  resident transpose for dgrad; last Ra columns hold A_j^T
  A_stack [64, H]       LoRA down-projection operand (rows j*r..), zero padded to one 64-row TMA box
  Bblk [64, 3H]         block-diagonal (alpha/r)*B_j^T: g = dQKV . Bblk^T gives all LoRA mid-gradients in one GEMM
  Wo, WoT [H,H]; Wi [F,H], WiT [H,F]; Wo2 [H,F], Wo2T [F,H]; biases / LN params fp32
Activations: residual path fp32, GEMM operands bf16.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import ops
from .lora import LoraBank

bf16, f32 = torch.bfloat16, torch.float32


def _aug_buf(rows: int, cols: int, ra: int, device, zero: bool = False) -> torch.Tensor:
    """[rows, cols+ra] bf16 view whose row stride is padded to cols+64 when ra > 0, so that every row starts on a
    128-byte boundary (TMA 128B-swizzled boxes then touch aligned lines; measured +20 % on the K-augmented GEMMs)"""
    ld = cols + (64 if ra else 0)
    base = (torch.zeros if zero else torch.empty)(rows, ld, dtype=bf16, device=device)
    return base[:, :cols + ra]


class _Ctx:
    """activations of one forward call kept for its backward"""
    pass


class BertEncoder(torch.nn.Module):
    LORA_TARGETS = ("query", "key", "value")          # reference rag_e2e_base_model.py:66-68

    def __init__(self, cfg: Dict, state_dict: Dict[str, torch.Tensor], device="cuda", lora: bool = False,
                 lora_seed: int = 0):
        super().__init__()
        self.cfg = cfg
        self.H = H = cfg["hidden_size"]
        self.F = F = cfg["intermediate_size"]
        self.nl = cfg["num_hidden_layers"]
        self.nh = cfg["num_attention_heads"]
        self.hd = H // self.nh
        self.V = cfg["vocab_size"]
        self.eps = float(cfg.get("layer_norm_eps", 1e-12))
        self.dev = torch.device(device)
        if self.hd not in (32, 64, 128):
            raise NotImplementedError(f"head_dim {self.hd} not supported by the attention kernels")
        self.r = 8
        self.Ra = 3 * self.r if lora else 0
        sd = {k[len("bert."):] if k.startswith("bert.") else k: v for k, v in state_dict.items()}
        g = lambda k, dt: sd[k].to(device=self.dev, dtype=dt).contiguous()
        self.word = g("embeddings.word_embeddings.weight", bf16)
        self.pos = g("embeddings.position_embeddings.weight", bf16)
        self.type0 = g("embeddings.token_type_embeddings.weight", bf16)[0].contiguous()
        self.emb_g = g("embeddings.LayerNorm.weight", f32)
        self.emb_b = g("embeddings.LayerNorm.bias", f32)
        self.layers: List[Dict[str, torch.Tensor]] = []
        for l in range(self.nl):
            p = f"encoder.layer.{l}."
            W = {}
            wq, wk, wv = (g(p + f"attention.self.{n}.weight", bf16) for n in self.LORA_TARGETS)
            W["Wqkv_aug"] = _aug_buf(3 * H, H, self.Ra, self.dev, zero=True)
            W["Wqkv_aug"][:, :H] = torch.cat([wq, wk, wv], 0)
            W["WqkvT_aug"] = _aug_buf(H, 3 * H, self.Ra, self.dev, zero=True)
            W["WqkvT_aug"][:, :3 * H] = torch.cat([wq, wk, wv], 0).t()
            W["bqkv"] = torch.cat([g(p + f"attention.self.{n}.bias", f32) for n in self.LORA_TARGETS])
            if lora:
                W["A_stack"] = torch.zeros(64, H, dtype=bf16, device=self.dev)
                W["Bblk"] = torch.zeros(64, 3 * H, dtype=bf16, device=self.dev)
            W["Wo"] = g(p + "attention.output.dense.weight", bf16)
            W["WoT"] = W["Wo"].t().contiguous()
            W["bo"] = g(p + "attention.output.dense.bias", f32)
            W["ln1_g"] = g(p + "attention.output.LayerNorm.weight", f32)
            W["ln1_b"] = g(p + "attention.output.LayerNorm.bias", f32)
            W["Wi"] = g(p + "intermediate.dense.weight", bf16)
            W["WiT"] = W["Wi"].t().contiguous()
            W["bi"] = g(p + "intermediate.dense.bias", f32)
            W["Wo2"] = g(p + "output.dense.weight", bf16)
            W["Wo2T"] = W["Wo2"].t().contiguous()
            W["bo2"] = g(p + "output.dense.bias", f32)
            W["ln2_g"] = g(p + "output.LayerNorm.weight", f32)
            W["ln2_b"] = g(p + "output.LayerNorm.bias", f32)
            self.layers.append(W)
        self.pooler = {k: v for k, v in sd.items() if k.startswith("pooler.")}   # carried for save_pretrained only
        self.lora: Optional[LoraBank] = None
        if lora:
            specs = [(f"encoder.layer.{l}.attention.self.{n}", H, H) for l in range(self.nl) for n in self.LORA_TARGETS]
            self.lora = LoraBank(specs, r=self.r, alpha=16, dropout=0.05, device=self.dev, seed=lora_seed)
            self.lora_flat = torch.nn.Parameter(self.lora.flat, requires_grad=True)
            self.lora_flat.grad = self.lora.grad
            self.repack_lora()

    # ------------------------------------------------------------------------------------------------------------
    def _pack_entries(self):
        H, r, s = self.H, self.r, self.lora.scale
        for l, W in enumerate(self.layers):
            for j, n in enumerate(self.LORA_TARGETS):
                name = f"encoder.layer.{l}.attention.self.{n}"
                A, B = self.lora.A[name], self.lora.B[name]             # [r,H], [H,r] fp32
                yield (B, r, 1, W["Wqkv_aug"][j * H:(j + 1) * H, H + j * r:], H, r, s)      # (alpha/r) * B
                yield (A, 1, H, W["WqkvT_aug"][:, 3 * H + j * r:], H, r, 1.0)                # A^T
                yield (A, H, 1, W["A_stack"][j * r:(j + 1) * r], r, H, 1.0)                  # A
                yield (B, 1, r, W["Bblk"][j * r:(j + 1) * r, j * H:], r, H, s)               # (alpha/r) * B^T, block j

    def repack_lora(self) -> None:
        """refresh the bf16 LoRA blocks inside the augmented weights from the fp32 master copies: ONE launch over a
        device-resident table of (source, strides, destination) records built once (pointers never move)"""
        if self.lora is None:
            return
        if getattr(self, "_pack_tab", None) is None:
            self._pack_tab = ops.build_pack_table(list(self._pack_entries()), self.dev)
        ops.pack_table_(self._pack_tab)

    # ------------------------------------------------------------------------------------------------------------
    def forward_hidden(self, ids: torch.Tensor, mask: torch.Tensor, save: bool = True):
        """ids, mask: int64 [B,L] on device -> (hidden fp32 [B,L,H], ctx)"""
        B, L = ids.shape
        M, H, F, Ra = B * L, self.H, self.F, self.Ra
        ctx = _Ctx()
        ctx.B, ctx.L, ctx.mask = B, L, mask.contiguous()
        ctx.layers = []
        z = ops.bert_embed(ids, self.word, self.pos, self.type0)
        x_aug = _aug_buf(M, H, Ra, self.dev)
        x32, _, mean, rstd = ops.layernorm_fwd(z, self.emb_g, self.emb_b, self.eps, y16=x_aug[:, :H])
        for W in self.layers:
            a = _Ctx()
            a.x_aug = x_aug
            if Ra:
                ops.skinny_gemm(x_aug[:, :H], W["A_stack"], x_aug[:, H:], K=H, R=Ra)             # u = x A^T  [M,3r]
            qkv = ops.gemm(x_aug, W["Wqkv_aug"], bias=W["bqkv"])                                   # [M,3H] (+LoRA via K-aug)
            att, lse = ops.attention_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], ctx.mask, B, L, self.nh, self.nh,
                                         self.hd, causal=False)
            z1 = ops.gemm(att, W["Wo"], out_dtype=f32, bias=W["bo"], resid=x32)                    # dense + residual
            h_aug = torch.empty(M, H, dtype=bf16, device=self.dev)
            h32, _, m1, r1 = ops.layernorm_fwd(z1, W["ln1_g"], W["ln1_b"], self.eps, y16=h_aug)
            pre = ops.gemm(h_aug, W["Wi"], bias=W["bi"])                                           # [M,F] pre-activation
            act = ops.gelu_fwd(pre)
            z2 = ops.gemm(act, W["Wo2"], out_dtype=f32, bias=W["bo2"], resid=h32)
            x_aug = _aug_buf(M, H, Ra, self.dev)
            x32, _, m2, r2 = ops.layernorm_fwd(z2, W["ln2_g"], W["ln2_b"], self.eps, y16=x_aug[:, :H])
            if save:
                a.qkv, a.att, a.lse, a.z1, a.m1, a.r1, a.h_aug, a.pre, a.act, a.z2, a.m2, a.r2 = \
                    qkv, att, lse, z1, m1, r1, h_aug, pre, act, z2, m2, r2
                ctx.layers.append(a)
        return x32.view(B, L, H), ctx

    # ------------------------------------------------------------------------------------------------------------
    def backward_hidden(self, ctx: _Ctx, d_hidden: torch.Tensor) -> None:
        """d_hidden fp32 [B,L,H]; accumulates LoRA gradients into self.lora.grad (base weights are frozen: PEFT mode)."""
        if self.lora is None:
            return                                           # nothing trainable below the pooled output
        M, H = ctx.B * ctx.L, self.H
        last, Wl = ctx.layers[self.nl - 1], self.layers[self.nl - 1]
        last._pre = ops.layernorm_bwd(last.z2, Wl["ln2_g"], last.m2, last.r2, dy_f32=d_hidden.reshape(M, H).contiguous())
        self._bwd_from_ln2(ctx, self.nl - 1)

    def _bwd_from_ln2(self, ctx: _Ctx, l_start: int) -> None:
        """continue the backward at layer l_start whose LN2 input gradient has already been computed (stashed in _pre)"""
        B, L = ctx.B, ctx.L
        M, H, Ra, r = B * L, self.H, self.Ra, self.r
        for l in range(l_start, -1, -1):
            W, a = self.layers[l], ctx.layers[l]
            dz2_32, dz2_16 = a._pre
            del a._pre
            dact = ops.gemm(dz2_16, W["Wo2T"])
            ops.gelu_bwd_(a.pre, dact)
            dh_16 = ops.gemm(dact, W["WiT"])
            dz1_32, dz1_16 = ops.layernorm_bwd(a.z1, W["ln1_g"], a.m1, a.r1, dy_f32=dz2_32, dy_bf16=dh_16)
            datt = ops.gemm(dz1_16, W["WoT"])
            dqkv_aug = _aug_buf(M, 3 * H, Ra, self.dev)
            ops.attention_bwd(a.qkv[:, :H], a.qkv[:, H:2 * H], a.qkv[:, 2 * H:], ctx.mask, a.att, a.lse, datt, B, L,
                              self.nh, self.nh, self.hd, causal=False, dq=dqkv_aug[:, :H], dk=dqkv_aug[:, H:2 * H],
                              dv=dqkv_aug[:, 2 * H:3 * H])
            for j, n in enumerate(self.LORA_TARGETS):
                # g_j = dY_j (alpha/r) B_j : only the target's own column block is read
                ops.skinny_gemm(dqkv_aug[:, j * H:(j + 1) * H], W["Bblk"][j * r:(j + 1) * r, j * H:(j + 1) * H],
                                dqkv_aug[:, 3 * H + j * r:], K=H, R=r)
            names = [f"encoder.layer.{l}.attention.self.{n}" for n in self.LORA_TARGETS]
            # dA[rr,k] += sum_m g_j[m,rr] x[m,k]: q and k share one pass over x (16-row MMA tile), v takes a second
            ops.lora_wgrad_(a.x_aug[:, :H], dqkv_aug[:, 3 * H:], self.lora.gA[names[0]], H, 1, H, 2 * r, 1.0,
                            out1=self.lora.gA[names[1]])
            ops.lora_wgrad_(a.x_aug[:, :H], dqkv_aug[:, 3 * H + 2 * r:], self.lora.gA[names[2]], H, 1, H, r, 1.0)
            for j, name in enumerate(names):
                # dB[n,rr] += (alpha/r) * sum_m dY_j[m,n] u_j[m,rr]
                ops.lora_wgrad_(dqkv_aug[:, j * H:(j + 1) * H], a.x_aug[:, H + j * r:], self.lora.gB[name], 1, r, H, r,
                                self.lora.scale)
            if l == 0:
                return
            dx_16 = ops.gemm(dqkv_aug, W["WqkvT_aug"])
            p, Wp = ctx.layers[l - 1], self.layers[l - 1]
            p._pre = ops.layernorm_bwd(p.z2, Wp["ln2_g"], p.m2, p.r2, dy_f32=dz1_32, dy_bf16=dx_16)
