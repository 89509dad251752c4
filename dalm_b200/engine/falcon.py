"""Falcon decoder (Falcon-7B architecture: parallel attention + MLP, multi-query attention, LayerNorm, GELU, rotary,
tied lm_head) — FORWARD launch sequence over the C-ABI kernels.

Mirrors `self.generator_model(input_ids=..., attention_mask=...).logits` of the reference
(dalm/models/rag_e2e_base_model.py:104-106) through HF FalconForCausalLM (`trust_remote_code=True`, :54) for BASELINE
config 5. The reference's generator LoRA targets (`q_proj`, `v_proj`, rag_e2e_base_model.py:76-77) do not exist in
Falcon (its fused projection is `query_key_value`), so peft would refuse `--use-peft generator|both`; with
`--use-peft retriever` the generator is frozen and its gradient is never needed (the generator loss reaches the retriever
only through the doc log-prob term of the in-batch kernel). This engine therefore builds the forward only and raises for
adapter requests, exactly where peft would.

Per layer (bf16 weights): Wqkv [(nh+2)*hd, H] fused q|k|v (one KV head), Wd [H,H], W1 [4H,H], W2 [H,4H]; LayerNorm
gain/bias fp32. Residual stream fp32; x_out = x + attn(LN(x)) + mlp(LN(x)) (one LayerNorm feeds both branches).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .. import ops

bf16, f32 = torch.bfloat16, torch.float32


class FalconDecoder(torch.nn.Module):
    LORA_TARGETS = ()

    def __init__(self, cfg: Dict, state_dict: Dict[str, torch.Tensor], device="cuda", lora: bool = False, lora_seed: int = 1):
        super().__init__()
        if lora:
            raise ValueError("Target modules ['q_proj', 'v_proj'] not found in the base model (Falcon fuses them into "
                             "`query_key_value`): the reference's generator LoRA config cannot apply to Falcon; use "
                             "use_peft='retriever'")
        if cfg.get("new_decoder_architecture", False) or cfg.get("alibi", False) or not cfg.get("parallel_attn", True) \
                or not cfg.get("multi_query", True) or cfg.get("bias", False):
            raise NotImplementedError("only the Falcon-7B architecture variant (parallel_attn, multi_query, rotary, no bias) is built")
        self.cfg = cfg
        self.H = H = cfg["hidden_size"]
        self.nl = cfg["num_hidden_layers"]
        self.nh = cfg["num_attention_heads"]
        self.hd = H // self.nh
        self.V = cfg["vocab_size"]
        self.F = cfg.get("ffn_hidden_size") or 4 * H
        self.eps = float(cfg.get("layer_norm_epsilon", 1e-5))
        self.theta = float(cfg.get("rope_theta", 10000.0))
        self.dev = torch.device(device)
        if self.hd not in (32, 64, 128):
            raise NotImplementedError(f"head_dim {self.hd} not supported by the attention kernels")
        self.Nq, self.Nkv = self.nh * self.hd, self.hd
        sd = state_dict
        g = lambda k, dt: sd[k].to(device=self.dev, dtype=dt).contiguous()
        self.embed = g("transformer.word_embeddings.weight", bf16)
        self.Vp = (self.V + 7) // 8 * 8
        lm = g("lm_head.weight", bf16) if "lm_head.weight" in sd else self.embed          # tied
        if self.Vp != self.V:
            lm = torch.cat([lm, torch.zeros(self.Vp - self.V, H, dtype=bf16, device=self.dev)], 0)
        self.lm_head = lm
        self.lnf_g, self.lnf_b = g("transformer.ln_f.weight", f32), g("transformer.ln_f.bias", f32)
        self.layers: List[Dict[str, torch.Tensor]] = []
        for l in range(self.nl):
            p = f"transformer.h.{l}."
            self.layers.append({
                "ln_g": g(p + "input_layernorm.weight", f32), "ln_b": g(p + "input_layernorm.bias", f32),
                "Wqkv": g(p + "self_attention.query_key_value.weight", bf16),
                "Wd": g(p + "self_attention.dense.weight", bf16),
                "W1": g(p + "mlp.dense_h_to_4h.weight", bf16), "W2": g(p + "mlp.dense_4h_to_h.weight", bf16),
            })
        self._rope_cache: Dict[int, tuple] = {}
        self.lora = None
        self.full = None
        self.trainable = False                                 # forward-only: no backward is built for Falcon yet
        self.drop_offset = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.eval()

    def repack_lora(self) -> None:
        pass

    def banks(self) -> list:
        return []

    def grad_buffers(self) -> list:
        return []

    def _rope(self, L: int):
        if L not in self._rope_cache:
            inv = 1.0 / (self.theta ** (torch.arange(0, self.hd, 2, dtype=torch.float32) / self.hd))
            fr = torch.outer(torch.arange(L, dtype=torch.float32), inv)
            self._rope_cache[L] = (fr.cos().to(self.dev).contiguous(), fr.sin().to(self.dev).contiguous())
        return self._rope_cache[L]

    def forward_logits(self, ids: torch.Tensor, mask: torch.Tensor, save: bool = False):
        """ids, mask int64 [B,L] -> (logits bf16 [B,L,V], None). Nothing is saved: the decoder is frozen."""
        B, L = ids.shape
        M, H = B * L, self.H
        cos_t, sin_t = self._rope(L)
        mask = mask.contiguous()
        x = ops.embed_gather(ids, self.embed)                                         # fp32 residual stream [M,H]
        for W in self.layers:
            _, h, _, _ = ops.layernorm_fwd(x, W["ln_g"], W["ln_b"], self.eps, want_f32=False)     # one LN feeds both branches
            qkv = ops.gemm(h, W["Wqkv"])                                              # [M, (nh+2)*hd]
            ops.rope_(qkv, 0, self.nh + 1, self.hd, cos_t, sin_t, L)                  # q heads then the single k head
            att, _ = ops.attention_fwd(qkv[:, :self.Nq], qkv[:, self.Nq:self.Nq + self.hd], qkv[:, self.Nq + self.hd:],
                                       mask, B, L, self.nh, 1, self.hd, causal=True)
            t = ops.gemm(att, W["Wd"], out_dtype=f32, resid=x)                        # x + attention branch
            h4 = ops.gemm(h, W["W1"], act=1)                                          # GELU(erf) fused in the epilogue
            x = ops.gemm(h4, W["W2"], out_dtype=f32, resid=t)                         # + MLP branch
        _, hf, _, _ = ops.layernorm_fwd(x, self.lnf_g, self.lnf_b, self.eps, want_f32=False)
        logits = ops.gemm(hf, self.lm_head)
        return logits.view(B, L, self.Vp)[:, :, :self.V], None

    def backward_logits(self, ctx, dlogits) -> None:
        return None                                                                   # frozen: nothing trainable
