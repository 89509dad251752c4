"""Greedy-decoding throughput of the generator (evaluation path, engine/decoding.py) at the reference's eval-rag sizes:
Llama-2-7B shape, query_batch_size 16 prompts (dalm/eval/eval_rag.py:76-80), prompt + answer inside max_length 256.
A decode step is HBM-bound: every weight is read once per token (12.95 GB of layer weights + 0.26 GB LM head in bf16) plus
the KV cache (2 * T * 4096 * 2 B per sequence and layer); `floor_ms` is that traffic at the measured HBM copy peak.
    python tools/bench_decode.py [B] [prompt_len] [max_length] [model]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dalm_b200 import _lib, synthetic
from dalm_b200.engine import params
from dalm_b200.engine.llama import LlamaDecoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L0 = int(sys.argv[2]) if len(sys.argv) > 2 else 192
T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
name = sys.argv[4] if len(sys.argv) > 4 else "Llama-2-7b-hf"
dev = torch.device("cuda:0")
cfg = synthetic.llama_config(name)
dec = LlamaDecoder(cfg, params.random_state_dict("llama", cfg, seed=0, dtype=torch.bfloat16, device=dev), device=dev)
torch.cuda.empty_cache()
g = torch.Generator().manual_seed(0)
ids = torch.randint(3, cfg["vocab_size"], (B, L0), generator=g).to(dev)
mask = torch.ones_like(ids)
mask[1, :17] = 0                                                   # one left-padded row, like a real batch


def timed(**kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    e0.record()
    out = dec.generate(input_ids=ids, attention_mask=mask, eos_token_id=[], pad_token_id=0, **kw)   # no EOS: fixed length
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out, _lib.launch_count()


from dalm_b200.engine import decoding

timed(max_new_tokens=6)                                            # warm-up (tensor maps, allocator)
pre_ms, _, _ = timed(max_new_tokens=1)                             # prefill + one argmax
steps = T - L0 - 1                                                 # decode steps after the prefill's token
os.environ["DALM_B200_DECODE_GRAPH"] = "0"                         # the launch sequence issued from Python, per token
eager_ms, out_e, launches = timed(max_length=T)
os.environ["DALM_B200_DECODE_GRAPH"] = "1"                         # default: captured once, replayed per token
tot_ms, out, _ = timed(max_length=T)
mode = dict(decoding.LAST_RUN)
step_ms = (tot_ms - pre_ms) / steps
nl, H, F, V = cfg["num_hidden_layers"], cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
w_bytes = nl * (4 * H * H + 3 * H * F) * 2 + V * H * 2
kv_bytes = nl * B * 2 * ((L0 + T) // 2) * H * 2                    # average cache length over the run
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
floor_ms = (w_bytes + kv_bytes) / (hbm * 1e9) * 1e3
print(json.dumps({"what": "greedy decode, KV cache", "model": name, "batch": B, "prompt_len": L0, "max_length": T,
                  "prefill_ms": pre_ms, "total_ms": tot_ms, "decode_ms_per_token_step": step_ms,
                  "tokens_per_s": B / (step_ms * 1e-3), "launch_mode": mode,
                  "eager_ms_per_token_step": (eager_ms - pre_ms) / steps, "eager_launches_per_step": launches / (steps + 1),
                  "tokens_equal_eager": bool(torch.equal(out, out_e)),
                  "weight_bytes": w_bytes, "kv_bytes_avg": kv_bytes, "floor_ms": floor_ms, "frac_of_hbm_floor": floor_ms / step_ms,
                  "peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                  "out_shape": list(out.shape)}))
