"""CPU oracle of the `use_bnb` weight transform: bitsandbytes NF4 quantise -> dequantise (blocksize 64, no double
quantisation — BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4", bnb_4bit_compute_dtype=bfloat16), reference
dalm/models/rag_e2e_base_model.py:136-142).

TEST INFRASTRUCTURE ONLY. bitsandbytes is absent offline and unpinned in the reference's pyproject: "parity unpinned". This
restates its published algorithm (QLoRA, Dettmers et al. 2023, and bitsandbytes' `quantize_4bit` / `dequantize_4bit`): the
checkpoint tensor is cast to fp16, flattened row-major, split into blocks of 64; absmax per block in fp32; x * (1/absmax)
mapped to the nearest of the 16 NF4 levels (boundary = midpoint, `>` goes up); dequantised value = level * absmax, rounded to
fp16 (the dtype recorded in the quant state), which the forward then casts to bf16 for the matmul."""
from __future__ import annotations

import numpy as np

NF4 = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0],
               dtype=np.float32)


def roundtrip(w: np.ndarray):
    """-> (dequantised fp32 array of w's shape, codes uint8 [n], absmax fp32 [ceil(n/64)])"""
    flat = np.asarray(w, dtype=np.float32).reshape(-1).astype(np.float16).astype(np.float32)
    n = flat.size
    pad = (-n) % 64
    x = np.concatenate([flat, np.zeros(pad, np.float32)]).reshape(-1, 64)
    absmax = np.abs(x).max(axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (np.float32(1.0) / absmax).astype(np.float32)
        scaled = (x * inv[:, None]).astype(np.float32)          # all-zero blocks give 0 * inf = nan: overwritten below
    bounds = (np.float32(0.5) * (NF4[:-1] + NF4[1:])).astype(np.float32)
    codes = (scaled[:, :, None] > bounds[None, None, :]).sum(axis=2).astype(np.uint8)
    codes[absmax == 0] = 7
    deq = (NF4[codes] * absmax[:, None]).astype(np.float32).astype(np.float16).astype(np.float32)
    return deq.reshape(-1)[:n].reshape(np.shape(w)), codes.reshape(-1)[:n], absmax
