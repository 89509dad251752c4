"""Offline fixtures: synthetic (Abstract, Question, Answer) data, tokenizers and random-init model directories.

There is no network in the build / GPU environment, so pretrained checkpoints and tokenizers of bge-* / Llama-2 /
Falcon cannot be fetched. This module writes HF-layout directories (config.json + tokenizer files [+ safetensors])
with the PUBLIC architecture shapes (SURVEY §8 model table) so that the drop-in wrappers can be pointed at them exactly
like at a hub name. Weights are seeded random-init (std 0.02), as BASELINE.json's configs prescribe.
"""
from __future__ import annotations

import csv
import json
import os
from typing import Dict, List, Optional

import numpy as np

# ---------------------------------------------------------------------------------------------------------------
# architecture shapes
# ---------------------------------------------------------------------------------------------------------------
BERT_SHAPES: Dict[str, Dict] = {
    "bge-tiny": dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128),
    "bge-small-en": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=12, intermediate_size=1536),
    "bge-large-en": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
}
LLAMA_SHAPES: Dict[str, Dict] = {
    "llama-tiny": dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
                       intermediate_size=256),
    "llama-hd128": dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
                        intermediate_size=512),
    "llama-mini": dict(hidden_size=512, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=4,
                       intermediate_size=1408),
    "Llama-2-7b-hf": dict(hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=32,
                          intermediate_size=11008),
}


FALCON_SHAPES: Dict[str, Dict] = {
    "falcon-tiny": dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2),
    "falcon-mini": dict(hidden_size=448, num_hidden_layers=2, num_attention_heads=7),          # 7 q heads x 64, one KV head
    "falcon-7b": dict(hidden_size=4544, num_hidden_layers=32, num_attention_heads=71),
}


def falcon_config(name: str, vocab_size: int = 65024) -> Dict:
    s = FALCON_SHAPES[name]
    return dict(
        architectures=["FalconForCausalLM"], model_type="falcon", vocab_size=vocab_size, alibi=False,
        new_decoder_architecture=False, multi_query=True, parallel_attn=True, bias=False, layer_norm_epsilon=1e-5,
        rope_theta=10000.0, hidden_dropout=0.0, attention_dropout=0.0, initializer_range=0.02, bos_token_id=11,
        eos_token_id=11, tie_word_embeddings=True, ffn_hidden_size=4 * s["hidden_size"], **s,
    )


def bert_config(name: str, vocab_size: int = 30522) -> Dict:
    s = BERT_SHAPES[name]
    return dict(
        architectures=["BertModel"], model_type="bert", vocab_size=vocab_size, max_position_embeddings=512,
        type_vocab_size=2, hidden_act="gelu", layer_norm_eps=1e-12, hidden_dropout_prob=0.1,
        attention_probs_dropout_prob=0.1, initializer_range=0.02, pad_token_id=0, position_embedding_type="absolute", **s,
    )


def llama_config(name: str, vocab_size: int = 32000) -> Dict:
    s = LLAMA_SHAPES[name]
    return dict(
        architectures=["LlamaForCausalLM"], model_type="llama", vocab_size=vocab_size, max_position_embeddings=4096,
        hidden_act="silu", rms_norm_eps=1e-5, rope_theta=10000.0, initializer_range=0.02, bos_token_id=1, eos_token_id=2,
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False, attention_dropout=0.0,
        head_dim=s["hidden_size"] // s["num_attention_heads"], **s,
    )


# ---------------------------------------------------------------------------------------------------------------
# synthetic text
# ---------------------------------------------------------------------------------------------------------------
_SYL = ["ka", "to", "mi", "ren", "sol", "va", "qu", "ex", "pli", "dor", "an", "be", "cu", "fi", "gra", "hy", "jo", "lu",
        "ne", "os", "pa", "ri", "su", "ty", "ul", "vo", "wi", "xa", "yo", "ze"]


def word_list(n: int = 20000, seed: int = 7) -> List[str]:
    rng = np.random.default_rng(seed)
    words, seen = [], set()
    while len(words) < n:
        k = int(rng.integers(1, 5))
        w = "".join(_SYL[int(i)] for i in rng.integers(0, len(_SYL), size=k))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


def synthetic_rows(n_rows: int, seed: int = 1234, full: bool = False, words: Optional[List[str]] = None):
    """Zipf(1.1) words; lengths passage~U[110,160], query~U[12,40], answer~U[3,25] words.
    full=True: passage>=200, query>=60, answer>=150 words so every tokenised sequence hits truncation (masks all ones)."""
    words = words or word_list()
    rng = np.random.default_rng(seed)
    nw = len(words)

    def sample(k: int) -> str:
        idx = np.minimum(rng.zipf(1.1, size=k) - 1, nw - 1)
        return " ".join(words[int(i)] for i in idx)

    for _ in range(n_rows):
        if full:
            lp, lq, la = int(rng.integers(200, 240)), int(rng.integers(60, 80)), int(rng.integers(150, 180))
        else:
            lp, lq, la = int(rng.integers(110, 161)), int(rng.integers(12, 41)), int(rng.integers(3, 26))
        yield {"Abstract": sample(lp), "Question": sample(lq), "Answer": sample(la)}


def write_csv(path: str, n_rows: int, seed: int = 1234, full: bool = False) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["Abstract", "Question", "Answer"])
        w.writeheader()
        for row in synthetic_rows(n_rows, seed=seed, full=full):
            w.writerow(row)
    return path


# ---------------------------------------------------------------------------------------------------------------
# tokenizers (trained offline on the synthetic word list with the `tokenizers` library)
# ---------------------------------------------------------------------------------------------------------------
def _corpus(n: int = 3000) -> List[str]:
    rows = list(synthetic_rows(n, seed=99))
    out = []
    for r in rows:
        out += [f"#query# {r['Question']}", f"#passage# {r['Abstract']}", f"#answer# {r['Answer']}"]
    return out


def build_bert_tokenizer(out_dir: str, vocab_size: int = 30522) -> str:
    """WordPiece vocabulary trained on the synthetic corpus, wrapped in transformers' BertTokenizer (lower-casing,
    [CLS]/[SEP], token_type_ids) like BAAI/bge-*'s tokenizer."""
    from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, trainers
    from transformers import BertTokenizer

    tok = Tokenizer(models.WordPiece(unk_token="[UNK]"))
    tok.normalizer = normalizers.BertNormalizer(lowercase=True)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    special = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    trainer = trainers.WordPieceTrainer(vocab_size=min(vocab_size, 8000), special_tokens=special, show_progress=False)
    tok.train_from_iterator(_corpus(), trainer)
    vocab = dict(tok.get_vocab())
    nxt = len(vocab)
    while nxt < vocab_size:          # pad up to the architecture's vocab size so ids span the real embedding table
        vocab[f"[unused{nxt}]"] = nxt
        nxt += 1
    bt = BertTokenizer(vocab=vocab, do_lower_case=True, model_max_length=512)
    os.makedirs(out_dir, exist_ok=True)
    bt.save_pretrained(out_dir)
    return out_dir


def build_llama_tokenizer(out_dir: str, vocab_size: int = 32000) -> str:
    """Llama-style BPE (metaspace, byte fallback, <unk>/<s>/</s> = 0/1/2) wrapped in transformers' LlamaTokenizer so that
    `add_eos_token = True` (reference train_rage2e.py:304) behaves as it does for the real Llama-2 tokenizer."""
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers
    from transformers import LlamaTokenizer

    tok = Tokenizer(models.BPE(unk_token="<unk>", fuse_unk=True, byte_fallback=True))
    tok.pre_tokenizer = pre_tokenizers.Metaspace(replacement="▁", prepend_scheme="first", split=False)
    byte_tokens = [f"<0x{i:02X}>" for i in range(256)]
    trainer = trainers.BpeTrainer(vocab_size=min(vocab_size, 6000), special_tokens=["<unk>", "<s>", "</s>"] + byte_tokens, show_progress=False)
    tok.train_from_iterator(_corpus(), trainer)
    vocab = tok.get_vocab()
    model_json = json.loads(tok.to_str())["model"]
    merges = [tuple(m) if isinstance(m, list) else tuple(m.split(" ")) for m in model_json["merges"]]
    nxt = len(vocab)
    while nxt < vocab_size:
        vocab[f"<extra_{nxt}>"] = nxt
        nxt += 1
    lt = LlamaTokenizer(vocab=vocab, merges=merges)
    lt.add_bos_token = True           # Llama-2 prepends <s>; baked into the saved post-processor
    os.makedirs(out_dir, exist_ok=True)
    lt.save_pretrained(out_dir)
    return out_dir


# ---------------------------------------------------------------------------------------------------------------
# model directories
# ---------------------------------------------------------------------------------------------------------------
def write_model_dir(out_dir: str, kind: str, name: str, vocab_size: Optional[int] = None, with_weights: bool = True,
                    seed: int = 0) -> str:
    """kind: 'bert' | 'llama'. Writes config.json, tokenizer files and (optionally) seeded random-init safetensors in HF
    parameter naming so both transformers (oracle) and dalm_b200 (product) can load the same directory."""
    os.makedirs(out_dir, exist_ok=True)
    if kind == "bert":
        cfg = bert_config(name, vocab_size or 30522)
        build_bert_tokenizer(out_dir, cfg["vocab_size"])
    elif kind == "llama":
        cfg = llama_config(name, vocab_size or 32000)
        build_llama_tokenizer(out_dir, cfg["vocab_size"])
    elif kind == "falcon":
        cfg = falcon_config(name, vocab_size or 65024)
        build_llama_tokenizer(out_dir, cfg["vocab_size"])       # any causal-LM tokenizer works for the synthetic fixture
    else:
        raise ValueError(kind)
    with open(os.path.join(out_dir, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    if not with_weights:
        with open(os.path.join(out_dir, "dalm_b200_random_init.json"), "w") as f:      # engine/params.py: random init at load time
            json.dump({"seed": seed}, f)
    if with_weights:
        import torch
        from safetensors.torch import save_file

        from .engine.params import random_state_dict

        sd = random_state_dict(kind, cfg, seed=seed, dtype=torch.float32, device="cpu")
        save_file({k: v.contiguous() for k, v in sd.items()}, os.path.join(out_dir, "model.safetensors"))
    return out_dir
