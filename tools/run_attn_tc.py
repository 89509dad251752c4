import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops
B, L, H, D = 18, 256, 32, 128
dev = torch.device("cuda:0")
qkv = torch.randn(B * L, 3 * H * D, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
mask = torch.ones(B, L, dtype=torch.int64, device=dev)
for _ in range(3):
    out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, H, H, D, True)
    do = torch.randn_like(out)
    ops.attention_tc_bwd(q, k, v, mask, out, lse, do, B, L, H, H, D, True)
torch.cuda.synchronize()
