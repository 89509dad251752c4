"""One eager cfg-3 (or cfg-2) training step between cudaProfilerStart / Stop, for `ncu --profile-from-start off`:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/profile_step.py [cfg-3|cfg-2]
Same models / batch / dropout settings as bench.py; two un-profiled warm-up steps first."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dalm_b200 import synthetic
from dalm_b200.engine import params
from dalm_b200.engine.bert import BertEncoder
from dalm_b200.engine.llama import LlamaDecoder
from dalm_b200.models.rag_e2e_base_model import AutoModelForRagE2E, Mode
from dalm_b200.optim import FusedAdam
from dalm_b200.training.utils import train_utils as tu

which = sys.argv[1] if len(sys.argv) > 1 else "cfg-3"
dev = torch.device("cuda:0")
bf = torch.bfloat16
bcfg = dict(synthetic.bert_config("bge-large-en"), _device_rng=True)
enc = BertEncoder(bcfg, params.random_state_dict("bert", bcfg, seed=0, dtype=bf, device=dev), device=dev, lora=True)
cfgd = bench.CONFIGS[which]
if which == "cfg-2":
    from dalm_b200.models.retriever_only_base_model import AutoModelForSentenceEmbedding
    model = AutoModelForSentenceEmbedding("", use_bnb=False, get_peft=True, _model=enc, _load_tokenizer=False)
    step_fn, repack = tu.fused_retriever_step, enc.repack_lora
else:
    lcfg = dict(synthetic.llama_config("Llama-2-7b-hf"), _device_rng=True)
    dec = LlamaDecoder(lcfg, params.random_state_dict("llama", lcfg, seed=0, dtype=bf, device=dev), device=dev, lora=True)
    model = AutoModelForRagE2E("", "", get_peft=Mode.BOTH, _retriever=enc, _generator=dec, _load_tokenizers=False)
    step_fn, repack = tu.fused_rag_step, model.repack
opt = FusedAdam(model.parameters(), lr=1e-4)
model.train()
tu._TWO_STREAMS = False                       # one stream: per-kernel durations are not inflated by cross-stream overlap
batches = [{k: v.to(dev) for k, v in b.items()} for b in bench.random_batches(3, cfgd, 0)]


def step(b):
    step_fn(model, b, 100.0, backward=True)
    opt.step(); repack(); opt.zero_grad()


for i in range(2):
    step(batches[i])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(batches[2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step of", which)
