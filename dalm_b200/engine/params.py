"""HF-named parameter dictionaries: seeded random init (no checkpoints are reachable offline) and directory loading."""
from __future__ import annotations

import json
import os
from typing import Dict

import torch


def _normal(gen: torch.Generator, shape, std: float, dtype, device) -> torch.Tensor:
    # CPU generator: reproducible across devices (tests, fixtures). A CUDA generator (bench-scale 7B models) draws on
    # the device directly.
    if gen.device.type == "cuda":
        return torch.empty(shape, dtype=torch.float32, device=gen.device).normal_(0.0, std, generator=gen).to(dtype)
    t = torch.empty(shape, dtype=torch.float32)
    t.normal_(mean=0.0, std=std, generator=gen)
    return t.to(device=device, dtype=dtype)


def random_state_dict(kind: str, cfg: Dict, seed: int = 0, dtype=torch.float32, device="cpu") -> Dict[str, torch.Tensor]:
    """HF parameter names for BertModel (no prefix) / LlamaForCausalLM, init N(0, initializer_range), LN = (1, 0)."""
    on_device = torch.device(device).type == "cuda" and cfg.get("_device_rng", False)
    gen = torch.Generator(device=device) if on_device else torch.Generator()
    gen.manual_seed(seed)
    std = float(cfg.get("initializer_range", 0.02))
    H = cfg["hidden_size"]
    sd: Dict[str, torch.Tensor] = {}
    ones = lambda n: torch.ones(n, dtype=dtype, device=device)
    zeros = lambda n: torch.zeros(n, dtype=dtype, device=device)
    if kind == "bert":
        F, V = cfg["intermediate_size"], cfg["vocab_size"]
        sd["embeddings.word_embeddings.weight"] = _normal(gen, (V, H), std, dtype, device)
        sd["embeddings.position_embeddings.weight"] = _normal(gen, (cfg["max_position_embeddings"], H), std, dtype, device)
        sd["embeddings.token_type_embeddings.weight"] = _normal(gen, (cfg["type_vocab_size"], H), std, dtype, device)
        sd["embeddings.LayerNorm.weight"] = ones(H)
        sd["embeddings.LayerNorm.bias"] = zeros(H)
        for l in range(cfg["num_hidden_layers"]):
            p = f"encoder.layer.{l}."
            for n in ("query", "key", "value"):
                sd[p + f"attention.self.{n}.weight"] = _normal(gen, (H, H), std, dtype, device)
                sd[p + f"attention.self.{n}.bias"] = _normal(gen, (H,), std, dtype, device)
            sd[p + "attention.output.dense.weight"] = _normal(gen, (H, H), std, dtype, device)
            sd[p + "attention.output.dense.bias"] = _normal(gen, (H,), std, dtype, device)
            sd[p + "attention.output.LayerNorm.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
            sd[p + "attention.output.LayerNorm.bias"] = _normal(gen, (H,), std, dtype, device)
            sd[p + "intermediate.dense.weight"] = _normal(gen, (F, H), std, dtype, device)
            sd[p + "intermediate.dense.bias"] = _normal(gen, (F,), std, dtype, device)
            sd[p + "output.dense.weight"] = _normal(gen, (H, F), std, dtype, device)
            sd[p + "output.dense.bias"] = _normal(gen, (H,), std, dtype, device)
            sd[p + "output.LayerNorm.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
            sd[p + "output.LayerNorm.bias"] = _normal(gen, (H,), std, dtype, device)
        sd["pooler.dense.weight"] = _normal(gen, (H, H), std, dtype, device)   # loaded by AutoModel, unused by the path
        sd["pooler.dense.bias"] = zeros(H)
    elif kind == "llama":
        F, V = cfg["intermediate_size"], cfg["vocab_size"]
        nh, nkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads", cfg["num_attention_heads"])
        hd = cfg.get("head_dim", H // nh)
        sd["model.embed_tokens.weight"] = _normal(gen, (V, H), std, dtype, device)
        for l in range(cfg["num_hidden_layers"]):
            p = f"model.layers.{l}."
            sd[p + "self_attn.q_proj.weight"] = _normal(gen, (nh * hd, H), std, dtype, device)
            sd[p + "self_attn.k_proj.weight"] = _normal(gen, (nkv * hd, H), std, dtype, device)
            sd[p + "self_attn.v_proj.weight"] = _normal(gen, (nkv * hd, H), std, dtype, device)
            sd[p + "self_attn.o_proj.weight"] = _normal(gen, (H, nh * hd), std, dtype, device)
            sd[p + "mlp.gate_proj.weight"] = _normal(gen, (F, H), std, dtype, device)
            sd[p + "mlp.up_proj.weight"] = _normal(gen, (F, H), std, dtype, device)
            sd[p + "mlp.down_proj.weight"] = _normal(gen, (H, F), std, dtype, device)
            sd[p + "input_layernorm.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
            sd[p + "post_attention_layernorm.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
        sd["model.norm.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
        sd["lm_head.weight"] = _normal(gen, (V, H), std, dtype, device)
    elif kind == "falcon":
        V = cfg["vocab_size"]
        nh = cfg["num_attention_heads"]
        hd = H // nh
        F = cfg.get("ffn_hidden_size") or 4 * H
        sd["transformer.word_embeddings.weight"] = _normal(gen, (V, H), std, dtype, device)
        for l in range(cfg["num_hidden_layers"]):
            p = f"transformer.h.{l}."
            sd[p + "input_layernorm.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
            sd[p + "input_layernorm.bias"] = _normal(gen, (H,), std, dtype, device)
            sd[p + "self_attention.query_key_value.weight"] = _normal(gen, ((nh + 2) * hd, H), std, dtype, device)
            sd[p + "self_attention.dense.weight"] = _normal(gen, (H, H), std, dtype, device)
            sd[p + "mlp.dense_h_to_4h.weight"] = _normal(gen, (F, H), std, dtype, device)
            sd[p + "mlp.dense_4h_to_h.weight"] = _normal(gen, (H, F), std, dtype, device)
        sd["transformer.ln_f.weight"] = ones(H) + _normal(gen, (H,), std, dtype, device)
        sd["transformer.ln_f.bias"] = _normal(gen, (H,), std, dtype, device)
        # lm_head is tied to the word embeddings (tie_word_embeddings=True): not stored separately
    else:
        raise ValueError(f"unknown model kind {kind!r}")
    return sd


def load_config(path: str) -> Dict:
    with open(os.path.join(path, "config.json")) as f:
        return json.load(f)


RANDOM_INIT_MARKER = "dalm_b200_random_init.json"


def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """Reads *.safetensors (single or sharded) or pytorch_model.bin from an HF-layout directory. A directory written by
    `synthetic.write_model_dir(..., with_weights=False)` carries a marker instead of weights ({"seed": n}): the public
    architecture is then random-initialised on the fly (on the GPU when there is one) - the offline stand-in for a hub
    checkpoint that BASELINE.json's configs prescribe ("random-init"), without writing 27 GB of Llama-2-7B to disk."""
    from safetensors.torch import load_file

    marker = os.path.join(path, RANDOM_INIT_MARKER)
    if os.path.exists(marker):
        with open(marker) as f:
            seed = int(json.load(f).get("seed", 0))
        cfg = load_config(path)
        on_gpu = torch.cuda.is_available()
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))) if on_gpu else torch.device("cpu")
        return random_state_dict(model_kind(cfg), dict(cfg, _device_rng=on_gpu), seed=seed,
                                 dtype=torch.bfloat16 if on_gpu else torch.float32, device=dev)
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors") and not f.startswith("adapter"))
    sd: Dict[str, torch.Tensor] = {}
    if files:
        for f in files:
            sd.update(load_file(os.path.join(path, f)))
        return sd
    binf = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(binf):
        return torch.load(binf, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no weights (model.safetensors / pytorch_model.bin) under {path}")


def model_kind(cfg: Dict) -> str:
    mt = cfg.get("model_type", "")
    if mt == "bert":
        return "bert"
    if mt == "llama":
        return "llama"
    if mt == "falcon":
        return "falcon"
    raise NotImplementedError(
        f"model_type {mt!r} is not built in dalm_b200 (supported: bert encoders; llama and falcon decoders); see DESIGN.md")


def is_bnb_linear_weight(name: str) -> bool:
    """the tensors `load_in_4bit` replaces: every nn.Linear weight except the LM head (transformers keeps `lm_head` and tied
    output embeddings out of the conversion); embeddings and norms are not Linear"""
    return (name.endswith(".weight") and "embed" not in name and "LayerNorm" not in name and "layernorm" not in name
            and "norm.weight" not in name and "ln_f" not in name and not name.startswith("lm_head"))


def bnb_nf4_state_dict(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """`use_bnb` (reference rag_e2e_base_model.py:136-142): fp32 tensors on `device` holding exactly the values the
    reference's 4-bit model computes with — nn.Linear weights through the NF4 quantise/dequantise round trip (csrc/nf4.cu),
    everything else through the fp16 cast transformers applies when a bitsandbytes config is given without torch_dtype."""
    from .. import ops

    out = {}
    for k, v in sd.items():
        t = v.to(device=device, dtype=torch.float32).contiguous().clone()
        if t.dim() == 2 and is_bnb_linear_weight(k):
            ops.nf4_roundtrip_(t)
        else:
            t = t.to(torch.float16).to(torch.float32)            # dtype casts only
        out[k] = t
    return out
