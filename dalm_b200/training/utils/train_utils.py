"""Loss functions and checkpoint hooks of the training step — same names and argument meaning as the reference's
dalm/training/utils/train_utils.py:12-138, executed by the dalm_b200 CUDA kernels.

Two ways in:
  * the stand-alone functions (`get_cosine_sim`, `get_nt_xent_loss`, `compute_marginalized_loss_from_logits`, ...)
    are differentiable through torch.autograd so a caller can keep the reference's loop body verbatim;
  * `fused_rag_step` / `fused_retriever_step` run the whole loop body (reference train_rage2e.py:431-471,
    train_retriever_only.py:365-376) as one launch sequence without autograd — what dalm_b200's own trainers use.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional

import torch

from ... import ops

f32, bf16, i64 = torch.float32, torch.bfloat16, torch.int64


# ----------------------------------------------------------------------------------------------------------------
# stand-alone differentiable functions
# ----------------------------------------------------------------------------------------------------------------
class _CosineSimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p, scale):
        ctx.save_for_backward(q, p)
        ctx.scale = float(scale)
        return ops.small_matmul(q, p, trans_b=True, alpha=float(scale))

    @staticmethod
    def backward(ctx, dS):
        q, p = ctx.saved_tensors
        dS = dS.contiguous().float()
        dq = ops.small_matmul(dS, p, alpha=ctx.scale)                    # dQ = s * dS P
        dp = ops.small_matmul(dS, q, trans_a=True, alpha=ctx.scale)      # dP = s * dS^T Q
        return dq, dp, None


def get_cosine_sim(query_embs: torch.Tensor, passage_embs: torch.Tensor, logit_scale: int) -> torch.Tensor:
    """reference :76-77"""
    return _CosineSimFn.apply(query_embs.float().contiguous(), passage_embs.float().contiguous(), logit_scale)


def _ce_square(scores: torch.Tensor, weights: torch.Tensor, nsum: torch.Tensor):
    """log_softmax(scores,1).diag() and d/dscores of  -(sum_i w_i * lsm_ii)/nsum  via the vocabulary-CE kernel:
    the [n,n] matrix is viewed as one sequence of n 'positions' over a vocabulary of n (+1 padding row)."""
    n = scores.shape[0]
    dev = scores.device
    lg = torch.zeros(1, n + 1, n, dtype=f32, device=dev)
    lg[0, :n] = scores
    ids = torch.zeros(1, n + 1, dtype=i64, device=dev)
    ids[0, 1:] = torch.arange(n, device=dev)
    w = torch.zeros(1, n + 1, dtype=i64, device=dev)
    w[0, 1:] = weights
    tok_lp, dl = ops.ce_marginal(lg, ids, w, nsum, need_grad=True)
    return tok_lp[0, :n], dl[0, :n], w


class _NtXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores):
        n = scores.shape[0]
        nsum = torch.full((1,), float(n), dtype=f32, device=scores.device)
        diag_lp, dS, w = _ce_square(scores.float().contiguous(), torch.ones(n, dtype=i64, device=scores.device), nsum)
        out = ops.finalize_loss(torch.cat([diag_lp, diag_lp.new_zeros(1)]).view(1, n + 1), w, nsum, None)
        ctx.save_for_backward(dS)
        return out[1].clone()

    @staticmethod
    def backward(ctx, g):
        (dS,) = ctx.saved_tensors
        return dS * g


def get_nt_xent_loss(sim_scores: torch.Tensor) -> torch.Tensor:
    """reference :80-88 — cross_entropy(sim, arange(n)), mean-reduced"""
    return _NtXentFn.apply(sim_scores)


def get_nll(log_probs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """reference :91-93 (pure indexing: no arithmetic)"""
    return -torch.gather(log_probs, 2, labels.unsqueeze(2)).squeeze(-1)


def marginalize_log_probs(logprobs_logits: torch.Tensor, doc_logprobs: torch.Tensor,
                          query_token_length: torch.Tensor) -> torch.Tensor:
    """reference :96-110 (slice / broadcast-add / concat on one sample; kept for API completeness — the training path
    uses the closed form inside ce_marginal + inbatch kernels)"""
    q = int(query_token_length)
    head = logprobs_logits[: q - 1, :]
    tail = logprobs_logits[q - 1:, :] + doc_logprobs
    return torch.cat([head, tail], dim=0)


class _MarginalizedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, input_ids, attention_mask, scores, qlen):
        dev = logits.device
        ids = input_ids.to(dev, i64).contiguous()
        mask = attention_mask.to(dev, i64).contiguous()
        cvec, nsum = ops.marginal_counts(mask, qlen.to(dev, i64).contiguous())
        lg = logits if logits.dtype in (bf16, f32) else logits.float()
        tok_lp, dl = ops.ce_marginal(lg.contiguous(), ids, mask, nsum, need_grad=True)
        # doc term: weights c_b (integer token counts) over the same normaliser N
        diag_lp, dS, w = _ce_square(scores.float().contiguous(), cvec.to(i64), nsum)
        n = scores.shape[0]
        lm_tok = ops.finalize_loss(tok_lp, mask, nsum, None)[1]
        doc = ops.finalize_loss(torch.cat([diag_lp, diag_lp.new_zeros(1)]).view(1, n + 1), w, nsum, None)[1]
        ctx.save_for_backward(dl, dS)
        ctx.logits_dtype = logits.dtype
        return (lm_tok + doc).clone()

    @staticmethod
    def backward(ctx, g):
        dl, dS = ctx.saved_tensors
        return (dl * g.to(dl.dtype)).to(ctx.logits_dtype), None, None, dS * g, None


def compute_marginalized_loss_from_logits(logits: torch.Tensor, input_tensors: torch.Tensor,
                                          attention_mask: torch.Tensor, scores: torch.Tensor,
                                          query_token_length: torch.Tensor) -> torch.Tensor:
    """reference :113-138"""
    if logits.shape[0] != scores.shape[0] or logits.shape[0] != query_token_length.shape[0]:
        # zip(..., strict=True) in the reference (:127-129)
        raise ValueError("logits, scores and query_token_length must have the same batch size")
    return _MarginalizedLossFn.apply(logits, input_tensors, attention_mask, scores, query_token_length)


# ----------------------------------------------------------------------------------------------------------------
# fused loop bodies
# ----------------------------------------------------------------------------------------------------------------
def _encode_pair(enc, q_ids, q_mask, p_ids, p_mask, save: bool):
    """hidden states of the query and passage batches + a closure running the encoder backward. BERT encoders take both
    batches through the weights in ONE pass (segments); a causal-LM retriever runs them one after the other."""
    if hasattr(enc, "forward_segments"):
        (hq, hp), c = enc.forward_segments([(q_ids, q_mask), (p_ids, p_mask)], save=save)
        return hq, hp, (lambda dq, dp: enc.backward_segments(c, [dq, dp]))
    hq, cq = enc.forward_hidden(q_ids, q_mask, save=save)
    hp, cp = enc.forward_hidden(p_ids, p_mask, save=save)

    def bwd(dq, dp):
        enc.backward_hidden(cp, dp)
        enc.backward_hidden(cq, dq)
    return hq, hp, bwd


_TWO_STREAMS = os.environ.get("DALM_B200_TWO_STREAMS", "1") != "0"
_CHUNKED_HEAD = os.environ.get("DALM_B200_CHUNKED_HEAD", "1") != "0"      # 0: materialise the [B,L,V] logits (A/B switch)
_SIDE_STREAMS: Dict[int, torch.cuda.Stream] = {}


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[idx]


def _inbatch(q_emb, p_emb, logit_scale: float, cvec, nsum, need_grad: bool, grad_out: float):
    """the fused in-batch loss over this rank's rows (the reference's semantics), or - DALM_B200_CROSS_RANK_NEGATIVES=1 on a
    multi-rank job - over the all-gathered rows of every rank (training/utils/negatives.py)"""
    from . import negatives
    if negatives.active():
        import torch.distributed as dist
        return negatives.global_inbatch_loss(q_emb, p_emb, logit_scale, cvec, nsum, need_grad, grad_out, rank=dist.get_rank(),
                                             world=dist.get_world_size(), loss_fn=ops.inbatch_loss)
    return ops.inbatch_loss(q_emb, p_emb, float(logit_scale), cvec, nsum, need_grad=need_grad, grad_out=grad_out)


def _pool_masks(model, q_mask, p_mask, autoregressive: bool):
    from ...models.rag_e2e_base_model import pooling_mask
    return pooling_mask(q_mask, autoregressive).contiguous(), pooling_mask(p_mask, autoregressive).contiguous()


def fused_rag_step(rag_model, batch: Dict[str, torch.Tensor], logit_scale: float, backward: bool = True,
                   grad_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """reference train_rage2e.py:431-471 as one launch sequence. Gradients (LoRA) are ACCUMULATED into the banks.
    returns {"loss": 0-d fp32 tensor, "losses": [Lc, Lm, total, N], "S": [B,B]}"""
    enc, dec = rag_model.retriever_model, rag_model.generator_model
    dev = enc.dev
    g = lambda k: batch[k].to(dev, i64, non_blocking=True).contiguous()
    q_ids, q_mask = g("retriever_query_input_ids"), g("retriever_query_attention_mask")
    p_ids, p_mask = g("retriever_passage_input_ids"), g("retriever_passage_attention_mask")
    g_ids, g_mask, qlen = g("generator_input_input_ids"), g("generator_input_attention_mask"), g("query_passage_input_len")
    train_enc = backward and enc.trainable
    train_dec = backward and dec.trainable
    if enc.training or dec.training:                         # fresh dropout masks per step (also inside a graph replay)
        ops.bump_counter_(enc.drop_offset)
        ops.bump_counter_(dec.drop_offset)
    cvec, nsum = ops.marginal_counts(g_mask, qlen)
    L_p, L_q = p_ids.shape[1], q_ids.shape[1]

    def retriever_branch():
        # encoder forward (both batches), pooling, fused in-batch loss (+ dQ, dP), encoder backward: independent of the decoder
        hq, hp, enc_bwd = _encode_pair(enc, q_ids, q_mask, p_ids, p_mask, train_enc)
        q_pm, p_pm = _pool_masks(rag_model, q_mask, p_mask, getattr(rag_model, "retriever_is_autoregressive", False))
        q_emb, q_norm = ops.pool_norm_fwd(hq, q_pm, rag_model.normalize)
        p_emb, p_norm = ops.pool_norm_fwd(hp, p_pm, rag_model.normalize)
        r = _inbatch(q_emb, p_emb, float(logit_scale), cvec, nsum, train_enc, grad_scale)
        if train_enc:
            enc_bwd(ops.pool_norm_bwd(q_emb, q_norm, r["dQ"], q_pm, L_q, rag_model.normalize),
                    ops.pool_norm_bwd(p_emb, p_norm, r["dP"], p_pm, L_p, rag_model.normalize))
        return r

    def generator_branch():
        if _CHUNKED_HEAD and hasattr(dec, "head_loss"):
            # lm_head + CE + the head's backward chunk by chunk over an L2-sized scratch: no [B,L,V] logits in HBM (engine/head.py)
            cg = dec.forward_final(g_ids, g_mask, save=train_dec)
            tok_lp, dhf = dec.head_loss(cg, g_ids, g_mask, nsum, need_grad=train_dec, grad_out=grad_scale)
            if train_dec:
                dec.backward_final(cg, dhf)
            return tok_lp
        logits, cg = dec.forward_logits(g_ids, g_mask, save=train_dec)
        tok_lp, dl = ops.ce_marginal(logits, g_ids, g_mask, nsum, need_grad=train_dec, inplace=True, grad_out=grad_scale)
        if train_dec:
            dec.backward_logits(cg, dl)
        return tok_lp

    # The two branches only meet in the scalar loss (the LM loss reaches the retriever through c_b / N, known up front),
    # so they run on two streams: the encoder's many small kernels fill the tails of the decoder's persistent GEMMs.
    if _TWO_STREAMS:
        main = torch.cuda.current_stream()
        side = _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            r = retriever_branch()
        tok_lp = generator_branch()
        main.wait_stream(side)
    else:
        r = retriever_branch()
        tok_lp = generator_branch()
    out = ops.finalize_loss(tok_lp, g_mask, nsum, r["losses"])
    return {"loss": out[2], "losses": out, "S": r["S"]}


def fused_retriever_step(model, batch: Dict[str, torch.Tensor], logit_scale: float, backward: bool = True,
                         grad_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """reference train_retriever_only.py:365-376"""
    enc = model.model
    dev = enc.dev
    g = lambda k: batch[k].to(dev, i64, non_blocking=True).contiguous()
    q_ids, q_mask, p_ids, p_mask = g("query_input_ids"), g("query_attention_mask"), g("passage_input_ids"), g("passage_attention_mask")
    train = backward and enc.trainable
    if enc.training:
        ops.bump_counter_(enc.drop_offset)
    hq, hp, enc_bwd = _encode_pair(enc, q_ids, q_mask, p_ids, p_mask, train)
    q_pm, p_pm = _pool_masks(model, q_mask, p_mask, getattr(model, "is_autoregressive", False))
    q_emb, q_norm = ops.pool_norm_fwd(hq, q_pm, model.normalize)
    p_emb, p_norm = ops.pool_norm_fwd(hp, p_pm, model.normalize)
    r = _inbatch(q_emb, p_emb, float(logit_scale), None, None, train, grad_scale)
    if train:
        enc_bwd(ops.pool_norm_bwd(q_emb, q_norm, r["dQ"], q_pm, q_ids.shape[1], model.normalize),
                ops.pool_norm_bwd(p_emb, p_norm, r["dP"], p_pm, p_ids.shape[1], model.normalize))
    return {"loss": r["losses"][0], "losses": r["losses"], "S": r["S"]}


class GraphedStep:
    """One CUDA graph for the whole forward+backward launch sequence of a fused step (~2 000 kernel launches at cfg-3):
    the host issues ONE graph launch per step instead of walking the Python launch sequence, so step time is
    independent of host speed. Inputs are copied into static device buffers; LoRA gradients accumulate into the banks'
    persistent buffers exactly as in eager mode. Batches whose shapes differ from the captured ones (a short final
    batch) run eagerly."""

    def __init__(self, step_fn, model, example_batch: Dict[str, torch.Tensor], logit_scale: float, grad_scale: float = 1.0,
                 zero_grads=None):
        dev = next(model.parameters()).device
        self.step_fn, self.model, self.logit_scale, self.grad_scale = step_fn, model, float(logit_scale), float(grad_scale)
        if self.grad_scale != 1.0:                              # gradient accumulation: the captured wgrad GEMMs must always +=
            for p in model.parameters():
                bank = getattr(p, "_dalm_bank", None)
                if bank is not None:
                    bank.force_accumulate = True
        self.static = {k: v.to(dev, i64).contiguous().clone() for k, v in example_batch.items() if torch.is_tensor(v)}
        self.shapes = {k: tuple(v.shape) for k, v in self.static.items()}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                           # warm-up off the capture stream: lazy tables, attributes
            for _ in range(2):
                step_fn(model, self.static, self.logit_scale, backward=True, grad_scale=self.grad_scale)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                                # the capture allocates from its own pool: do not keep the warm-up's
        if zero_grads is not None:                              # activations cached next to it (matters when HBM is nearly full)
            zero_grads()                                        # the warm-up accumulated gradients
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = step_fn(model, self.static, self.logit_scale, backward=True, grad_scale=self.grad_scale)
        if zero_grads is not None:
            zero_grads()

    def __call__(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if any(tuple(batch[k].shape) != shp for k, shp in self.shapes.items()):
            return self.step_fn(self.model, batch, self.logit_scale, backward=True, grad_scale=self.grad_scale)
        for k, buf in self.static.items():
            buf.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.out


# ----------------------------------------------------------------------------------------------------------------
# checkpoint hooks / adapter IO  (reference :12-73; PEFT adapter layout adapter_config.json + adapter_model.*)
# ----------------------------------------------------------------------------------------------------------------
ADAPTER_CONFIG = {
    "peft_type": "LORA", "r": 8, "lora_alpha": 16, "lora_dropout": 0.05, "bias": "none", "fan_in_fan_out": False,
    "inference_mode": False, "init_lora_weights": True,
}


def save_full_dir(engine_model, out_dir: str) -> None:
    """`save_pretrained` layout of a fully fine-tuned model (what the reference's hook writes for non-PEFT sub-models,
    train_utils.py:16-31): config.json + model.safetensors under HF parameter names, fp32"""
    from safetensors.torch import save_file

    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "config.json"), "w") as f:
        json.dump({k: v for k, v in engine_model.cfg.items() if not k.startswith("_")}, f, indent=1)
    save_file({k: v.contiguous() for k, v in engine_model.hf_state_dict().items()}, os.path.join(out_dir, "model.safetensors"))


def load_full_dir(engine_model, in_dir: str) -> None:
    from ...engine import params

    engine_model.load_hf_state_dict(params.load_state_dict(in_dir))


def save_adapter_dir(engine_model, out_dir: str, task_type: str, base_name: Optional[str] = None) -> None:
    os.makedirs(out_dir, exist_ok=True)
    if getattr(engine_model, "full", None) is not None:
        save_full_dir(engine_model, out_dir)
        return
    if engine_model.lora is None:
        return
    bank = engine_model.lora
    cfg = dict(ADAPTER_CONFIG, r=int(bank.r), lora_alpha=int(bank.alpha), lora_dropout=float(bank.dropout),
               task_type=task_type, target_modules=list(engine_model.LORA_TARGETS),
               base_model_name_or_path=base_name if base_name is not None else getattr(engine_model, "name_or_path", None))
    with open(os.path.join(out_dir, "adapter_config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    torch.save(engine_model.lora.peft_state_dict(), os.path.join(out_dir, "adapter_model.bin"))


def load_adapter_dir(engine_model, in_dir: str) -> None:
    if engine_model.lora is None:
        if not hasattr(engine_model, "enable_lora"):
            raise RuntimeError(f"{type(engine_model).__name__} cannot carry adapters")
        engine_model.enable_lora()          # PeftModel.from_pretrained(base, path) on a base built without get_peft
    path = os.path.join(in_dir, "adapter_model.bin")
    if os.path.exists(path):
        sd = torch.load(path, map_location="cpu", weights_only=True)
    else:
        from safetensors.torch import load_file

        sd = load_file(os.path.join(in_dir, "adapter_model.safetensors"))
    engine_model.lora.load_peft_state_dict(sd)
    engine_model.repack_lora()


def save_model_hook(models: List[torch.nn.Module], weights: List[Dict], output_dir: str) -> None:
    """reference :16-31: route wrapper weights to <dir>/retriever + <dir>/generator (RAG) or <dir> (sentence embedding)"""
    from ...models.rag_e2e_base_model import AutoModelForRagE2E
    from ...models.retriever_only_base_model import AutoModelForSentenceEmbedding

    for i, model in enumerate(models):
        if isinstance(model, AutoModelForSentenceEmbedding):
            save_adapter_dir(model.model, output_dir, "FEATURE_EXTRACTION")
        elif isinstance(model, AutoModelForRagE2E):
            save_adapter_dir(model.generator_model, os.path.join(output_dir, "generator"), "CAUSAL_LM")
            save_adapter_dir(model.retriever_model, os.path.join(output_dir, "retriever"), "FEATURE_EXTRACTION")
        else:
            raise NotImplementedError(f"Model type {type(model)} not supported")
        if weights:
            weights.pop()


def load_model_hook(models: List[torch.nn.Module], input_dir: str) -> None:
    """reference :34-73"""
    from ...models.rag_e2e_base_model import AutoModelForRagE2E
    from ...models.retriever_only_base_model import AutoModelForSentenceEmbedding

    while len(models) > 0:
        model = models.pop()
        if isinstance(model, AutoModelForRagE2E):
            for sub, name in ((model.generator_model, "generator"), (model.retriever_model, "retriever")):
                d = os.path.join(input_dir, name)
                if sub.lora is not None and os.path.exists(os.path.join(d, "adapter_config.json")):
                    load_adapter_dir(sub, d)
                elif getattr(sub, "full", None) is not None and os.path.exists(os.path.join(d, "model.safetensors")):
                    load_full_dir(sub, d)
        elif isinstance(model, AutoModelForSentenceEmbedding):
            if model.model.lora is not None and os.path.exists(os.path.join(input_dir, "adapter_config.json")):
                load_adapter_dir(model.model, input_dir)
            elif getattr(model.model, "full", None) is not None and os.path.exists(os.path.join(input_dir, "model.safetensors")):
                load_full_dir(model.model, input_dir)
        else:
            raise NotImplementedError(f"Model type {type(model)} not supported")
