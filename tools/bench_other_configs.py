"""Supplementary measurements of BASELINE.json's other configs (NOT the contract bench line; that is bench.py / cfg-3):
  cfg-2  train_retriever_only bge-large-en + PEFT, per-device bs=150, Lq 50 / Lp 128
  cfg-5  train_rage2e bge-large-en + Falcon-7B, bs=18, generator seq_len 2048, use_peft=retriever (frozen generator)
Same method as bench.py: synthetic 'full' rows, seeded random-init weights, train() mode (dropout on), one CUDA graph per step,
CUDA-event timing, warm-up, Adam + repack included.  python tools/bench_other_configs.py [cfg2|cfg5|cfg2full|cfg3full] [steps]
  cfg2full  cfg-2 without --use-peft: every BERT-large parameter trained (wgrad GEMMs, fp32 master + Adam over 335 M parameters)
  cfg3full  cfg-3 with use_peft=None (the reference's CLI default): bge-large AND Llama-2-7B fully fine-tuned"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import synthetic, _lib
from dalm_b200.engine import params
from dalm_b200.engine.bert import BertEncoder
from dalm_b200.optim import FusedAdam
from dalm_b200.training.utils.train_utils import GraphedStep, fused_rag_step, fused_retriever_step

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
bf16 = torch.bfloat16
g = torch.Generator().manual_seed(0)
bcfg = dict(synthetic.bert_config("bge-large-en"), _device_rng=True)
full = which in ("cfg2full", "cfg3full")
enc = BertEncoder(bcfg, params.random_state_dict("bert", bcfg, seed=0, dtype=bf16, device=dev), device=dev, lora=not full, full=full)


def rnd(B, L, V): return torch.randint(5, V, (B, L), generator=g)


def ones(B, L): return torch.ones(B, L, dtype=torch.int64)


if which in ("cfg2", "cfg2full"):
    from dalm_b200.models.retriever_only_base_model import AutoModelForSentenceEmbedding
    B = 150
    model = AutoModelForSentenceEmbedding("", use_bnb=False, get_peft=not full, _model=enc, _load_tokenizer=False)
    batches = [{"query_input_ids": rnd(B, 50, 30522), "query_attention_mask": ones(B, 50),
                "passage_input_ids": rnd(B, 128, 30522), "passage_attention_mask": ones(B, 128)} for _ in range(4)]
    step_fn, label = fused_retriever_step, "cfg-2 train_retriever_only bge-large-en + PEFT, bs=150, Lq50/Lp128"
    tflop = 33.1
    if full:
        label, tflop = "cfg-2 WITHOUT PEFT (full fine-tuning of bge-large-en), bs=150, Lq50/Lp128", 49.2
elif which == "cfg3full":
    from dalm_b200.engine.llama import LlamaDecoder
    from dalm_b200.models.rag_e2e_base_model import AutoModelForRagE2E
    B, LG = 18, 256
    lcfg = dict(synthetic.llama_config("Llama-2-7b-hf"), _device_rng=True)
    dec = LlamaDecoder(lcfg, params.random_state_dict("llama", lcfg, seed=1, dtype=bf16, device=dev), device=dev, full=True)
    torch.cuda.empty_cache()
    model = AutoModelForRagE2E("", "", get_peft=None, _retriever=enc, _generator=dec, _load_tokenizers=False)
    batches = [{"retriever_query_input_ids": rnd(B, 50, 30522), "retriever_query_attention_mask": ones(B, 50),
                "retriever_passage_input_ids": rnd(B, 128, 30522), "retriever_passage_attention_mask": ones(B, 128),
                "generator_input_input_ids": rnd(B, LG, 32000), "generator_input_attention_mask": ones(B, LG),
                "query_passage_input_len": torch.full((B,), 200)} for _ in range(4)]
    step_fn, label = fused_rag_step, "cfg-3 with use_peft=None (full fine-tuning of bge-large-en + Llama-2-7B), bs=18, Lq50/Lp128/Lg256"
    tflop = 190.4
else:
    from dalm_b200.engine.falcon import FalconDecoder
    from dalm_b200.models.rag_e2e_base_model import AutoModelForRagE2E, Mode
    B, LG = 18, 2048
    fcfg = dict(synthetic.falcon_config("falcon-7b"), _device_rng=True)
    gen_full = which == "cfg5full"        # the reference's semantics for --use-peft retriever: the generator is fully fine-tuned
    dec = FalconDecoder(fcfg, params.random_state_dict("falcon", fcfg, seed=0, dtype=bf16, device=dev), device=dev, full=gen_full)
    torch.cuda.empty_cache()
    model = AutoModelForRagE2E("", "", get_peft=Mode.RETRIEVER, _retriever=enc, _generator=dec, _load_tokenizers=False)
    batches = [{"retriever_query_input_ids": rnd(B, 50, 30522), "retriever_query_attention_mask": ones(B, 50),
                "retriever_passage_input_ids": rnd(B, 128, 30522), "retriever_passage_attention_mask": ones(B, 128),
                "generator_input_input_ids": rnd(B, LG, 65024), "generator_input_attention_mask": ones(B, LG),
                "query_passage_input_len": torch.full((B,), 700)} for _ in range(4)]
    step_fn, label = fused_rag_step, "cfg-5 train_rage2e bge-large-en + Falcon-7B (frozen, use_peft=retriever), bs=18, Lg=2048"
    tflop = 1.935 * 2 + 0.033 * 3 + 510.3 + 43.9        # encoder fwd+bwd (PEFT) + generator forward only (SURVEY §8d terms)
    if gen_full:
        label = "cfg-5 train_rage2e bge-large-en (LoRA) + Falcon-7B FULLY fine-tuned (reference semantics of use_peft=retriever), bs=18, Lg=2048, layer recomputation"
        tflop = 1.935 * 2 + 0.033 * 3 + 3 * (510.3 + 43.9)     # algorithmic full-FT work (3 x fwd); the recomputed forward is NOT counted

model.train()
opt = FusedAdam(model.parameters(), lr=1e-4)
dbs = [{k: v.to(dev) for k, v in b.items()} for b in batches]
graphed = GraphedStep(step_fn, model, dbs[0], 100.0, zero_grads=opt.zero_grad) if os.environ.get("NOGRAPH") != "1" else (lambda b: step_fn(model, b, 100.0, backward=True))


def one(i):
    out = graphed(dbs[i % len(dbs)])
    opt.step()
    if not full:
        (model.model if which == "cfg2" else model.retriever_model).repack_lora()
    opt.zero_grad()
    return out["loss"]


for i in range(3): one(i)
torch.cuda.synchronize()
_lib.reset_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps): loss = one(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"config": label, "samples_per_s": B / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "algorithmic_tflop_per_step": tflop,
                  "tflops": tflop / (ms * 1e-3), "loss_last": float(loss.item()), "dropout": "on", "launch": "CUDA graph", "n_gpus": 1,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
