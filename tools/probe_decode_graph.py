"""Separates the one-off CUDA-graph capture cost of `generate` from the per-token replay time: two forced-graph runs of
different lengths (the slope is the replayed step) next to an eager run (Llama-2-7B shape, 16 prompts x 64 tokens).
    python tools/probe_decode_graph.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import synthetic
from dalm_b200.engine import params
from dalm_b200.engine.llama import LlamaDecoder

dev = torch.device("cuda:0")
cfg = synthetic.llama_config("Llama-2-7b-hf")
dec = LlamaDecoder(cfg, params.random_state_dict("llama", cfg, seed=0, dtype=torch.bfloat16, device=dev), device=dev)
ids = torch.randint(3, 32000, (16, 64), device=dev)
mask = torch.ones_like(ids)


def run(T, mode):
    os.environ["DALM_B200_DECODE_GRAPH"] = mode
    torch.cuda.synchronize(); t = time.perf_counter()
    dec.generate(input_ids=ids, attention_mask=mask, eos_token_id=[], pad_token_id=0, max_length=T)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3


run(80, "1")
r = {"graph_192_ms": run(192, "1"), "graph_320_ms": run(320, "1"), "eager_192_ms": run(192, "0")}
r["replay_ms_per_step"] = (r["graph_320_ms"] - r["graph_192_ms"]) / 128
r["eager_ms_per_step"] = r["eager_192_ms"] / 128
r["capture_overhead_ms"] = r["graph_192_ms"] - 126 * r["replay_ms_per_step"] - 2 * r["eager_ms_per_step"]
print(json.dumps(r))
