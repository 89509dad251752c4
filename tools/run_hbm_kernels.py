"""Launches every HBM-bound training kernel once or twice at its cfg-3 / cfg-2 shape (for `ncu --metrics dram__bytes_*,
gpu__time_duration.sum`): RMSNorm / LayerNorm fwd+bwd, SwiGLU, RoPE, GELU, vocabulary CE, pool+normalise, Adam."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32
M, H, F, V, B, L = 4608, 4096, 11008, 32000, 18, 256
for rep in range(2):
    x = torch.randn(M, H, device=dev); g = torch.ones(H, device=dev)
    h, rstd = ops.rmsnorm_fwd(x, g, 1e-5)
    dh = torch.randn(M, H, device=dev).to(bf16); dres = torch.randn(M, H, device=dev)
    ops.rmsnorm_bwd(x, g, rstd, dh, dres_in=dres)
    gu = torch.randn(M, 2 * F, device=dev).to(bf16)
    act = ops.swiglu_fwd(gu, F)
    ops.swiglu_bwd_(gu, torch.randn(M, F, device=dev).to(bf16), F)
    qkv = torch.randn(M, 3 * H, device=dev).to(bf16)
    fr = torch.outer(torch.arange(L, dtype=f32), 1.0 / (10000 ** (torch.arange(0, 128, 2, dtype=f32) / 128)))
    ops.rope_(qkv, 0, 64, 128, fr.cos().to(dev), fr.sin().to(dev), L)
    logits = torch.randn(B, L, V, device=dev).to(bf16)
    ids = torch.randint(0, V, (B, L), device=dev); mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    nsum = torch.tensor([float(B * (L - 1))], device=dev)
    ops.ce_marginal(logits, ids, mask, nsum, inplace=True)
    # encoder shapes (cfg-2: 150 x (50 + 128) tokens x 1024)
    Me, He, Fe = 26700, 1024, 4096
    z = torch.randn(Me, He, device=dev); ge = torch.ones(He, device=dev); be = torch.zeros(He, device=dev)
    y32, y16, mean, rs = ops.layernorm_fwd(z, ge, be, 1e-12)
    ops.layernorm_bwd(z, ge, mean, rs, dy_f32=torch.randn(Me, He, device=dev), dy_bf16=torch.randn(Me, He, device=dev).to(bf16))
    pre = torch.randn(Me, Fe, device=dev).to(bf16)
    a = ops.gelu_fwd(pre); ops.gelu_bwd_(pre, torch.randn(Me, Fe, device=dev).to(bf16))
    hid = torch.randn(150, 128, He, device=dev); pm = torch.ones(150, 128, dtype=torch.int64, device=dev)
    emb, nrm = ops.pool_norm_fwd(hid, pm, True)
    ops.pool_norm_bwd(emb, nrm, torch.randn(150, He, device=dev), pm, 128, True)
    hid18 = torch.randn(18, 128, He, device=dev); pm18 = torch.ones(18, 128, dtype=torch.int64, device=dev)
    ops.pool_norm_fwd(hid18, pm18, True)
    n = 5_373_952                                                    # both LoRA banks
    pp = torch.randn(n, device=dev); gg = torch.randn(n, device=dev); m1 = torch.zeros(n, device=dev); v1 = torch.zeros(n, device=dev)
    ops.adam_step_(pp, gg, m1, v1, 1e-4, 0.9, 0.999, 1e-8, 1)
torch.cuda.synchronize()
print("done")
