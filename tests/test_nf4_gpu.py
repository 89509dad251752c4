"""-m gpu: `use_bnb` — the NF4 quantise/dequantise round trip (csrc/nf4.cu) bit-exact against the oracle restatement of
bitsandbytes' algorithm (oracle/nf4.py, "parity unpinned": bitsandbytes is absent offline), and the wrappers loading through it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,scale,seed", [((64,), 1.0, 0), ((5, 40), 0.02, 1), ((384, 384), 0.02, 2), ((1000, 129), 3.0, 3),
                                              ((7,), 1e-3, 4)])
def test_nf4_roundtrip_bit_exact_vs_oracle(cuda_dev, shape, scale, seed):
    from dalm_b200 import ops
    from oracle import nf4
    w = (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)
    if w.size > 200:
        w.reshape(-1)[64:128] = 0.0                                               # an all-zero block
        w.reshape(-1)[130] = w.reshape(-1)[128:192].__abs__().max() * 2            # a block dominated by one outlier
    want, codes, absmax = nf4.roundtrip(w)
    t = torch.from_numpy(w.copy()).to(cuda_dev)
    got, gc, ga = ops.nf4_roundtrip_(t, want_codes=True)
    assert np.array_equal(gc.cpu().numpy(), codes)                                 # integer codes: bit-exact
    assert np.array_equal(ga.cpu().numpy(), absmax)
    assert np.array_equal(got.cpu().numpy(), want)                                 # dequantised fp16 values: bit-exact
    again = ops.nf4_roundtrip_(got.clone())
    assert torch.equal(again, got)                                                 # size-independent property: idempotent
    err = np.abs(want - w.astype(np.float16).astype(np.float32)).reshape(-1)
    bound = np.repeat(absmax, 64)[: w.size] * 0.153 + 1e-6                         # half the widest gap between NF4 levels (0.3038)
    assert (err <= bound).all()


def test_nf4_boundaries_and_levels(cuda_dev):
    from dalm_b200 import ops
    from oracle import nf4
    levels = nf4.NF4.astype(np.float32)
    mids = (np.float32(0.5) * (levels[:-1] + levels[1:])).astype(np.float32)
    # a block whose absmax is exactly 1: levels map to themselves, midpoints go DOWN (`>` in bitsandbytes' decision tree)
    blk = np.zeros(64, np.float32); blk[:16] = levels; blk[16:31] = mids; blk[31] = np.nextafter(mids[3], np.float32(1))
    t = torch.from_numpy(blk.copy()).to(cuda_dev)
    _, codes, absmax = ops.nf4_roundtrip_(t, want_codes=True)
    _, ocodes, _ = nf4.roundtrip(blk)
    c = codes.cpu().numpy()
    assert np.array_equal(c, ocodes) and absmax.item() == 1.0
    assert c[:16].tolist() == list(range(16))


def test_wrappers_load_through_nf4(cuda_dev, tmp_path):
    """use_bnb=True (the reference's retriever-only default): embeddings equal the oracle HF model loaded with the oracle's NF4
    values, and differ from the un-quantised model; use_bnb without PEFT on the same sub-model is refused"""
    from safetensors.torch import load_file
    from dalm_b200 import synthetic
    from dalm_b200.engine import params
    from dalm_b200.models.rag_e2e_base_model import AutoModelForRagE2E, Mode, inference_only
    from dalm_b200.models.retriever_only_base_model import AutoModelForSentenceEmbedding
    from oracle import models as om, nf4
    import json, os
    rdir = synthetic.write_model_dir(str(tmp_path / "bge-tiny"), "bert", "bge-tiny", vocab_size=1200)
    gdir = synthetic.write_model_dir(str(tmp_path / "llama-tiny"), "llama", "llama-tiny", vocab_size=904)
    m_q = AutoModelForSentenceEmbedding(rdir, use_bnb=True, get_peft=True)
    m_f = AutoModelForSentenceEmbedding(rdir, use_bnb=False, get_peft=True)
    sd = load_file(os.path.join(rdir, "model.safetensors"))
    qsd = {}
    for k, v in sd.items():
        if v.dim() == 2 and params.is_bnb_linear_weight(k):
            qsd[k] = torch.from_numpy(nf4.roundtrip(v.float().numpy())[0])
        else:
            qsd[k] = v.half().float()
    n_quant = sum(v.dim() == 2 and params.is_bnb_linear_weight(k) for k, v in sd.items())
    assert n_quant == 2 * 6 + 1                                                    # 6 Linear per layer x 2 layers + pooler
    bert = om.build_bert(json.load(open(os.path.join(rdir, "config.json"))), qsd)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(5, 1200, (4, 24), generator=g); mask = torch.ones_like(ids); mask[1, 15:] = 0
    want = om.retrieval_forward(bert, ids, mask).detach()
    with torch.no_grad():
        got_q, got_f = m_q(ids, mask).cpu(), m_f(ids, mask).cpu()
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(got_q, want) < 2e-2                                                 # bf16 forward vs fp32 oracle on the SAME NF4 values
    assert rel(got_f, want) > 2 * rel(got_q, want)                                 # and the quantisation is really applied
    with pytest.raises(NotImplementedError):
        AutoModelForSentenceEmbedding(rdir, use_bnb=True, get_peft=False)          # 4-bit base + no adapters: nothing trainable
    rag = AutoModelForRagE2E(rdir, gdir, get_peft=Mode.BOTH, use_bnb=Mode.GENERATOR)
    gsd = load_file(os.path.join(gdir, "model.safetensors"))
    w0 = gsd["model.layers.0.mlp.down_proj.weight"].float()
    assert torch.equal(rag.generator_model.layers[0]["Wd"].cpu().float(), torch.from_numpy(nf4.roundtrip(w0.numpy())[0]).bfloat16().float())
    assert torch.equal(rag.generator_model.lm_head.cpu().float()[:904], gsd["lm_head.weight"].half().float().bfloat16().float())   # head: fp16 cast only
