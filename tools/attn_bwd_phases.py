"""clock64 timeline of CTA 0 of the pipelined dKdV kernel (first 8 halves / 4 items) at the cfg-3 decoder shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops, _lib
B, L, H, D = 18, 256, 32, 128
if len(sys.argv) > 1 and sys.argv[1] == "bge":
    B, L, H, D = 150, 128, 16, 64
dev = torch.device("cuda:0")
qkv = torch.randn(B * L, 3 * H * D, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
mask = torch.ones(B, L, dtype=torch.int64, device=dev)
causal = D == 128
out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, H, H, D, causal)
do = torch.randn_like(out)
for _ in range(3): ops.attention_tc_bwd(q, k, v, mask, out, lse, do, B, L, H, H, D, causal)
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
_lib.load().dalm_b200_attention_tc_set_debug(dbg.data_ptr())
ops.attention_tc_bwd(q, k, v, mask, out, lse, do, B, L, H, H, D, causal)
torch.cuda.synchronize()
_lib.load().dalm_b200_attention_tc_set_debug(None)
t = dbg.cpu().tolist()
names = {}
for n in range(8):
    names[3 * n] = f"MMA  half{n}: streamed tiles landed -> issue S/dP"
    names[3 * n + 1] = f"MMA  half{n}: p_ready"
    names[3 * n + 2] = f"MMA  half{n}: issue dV/dK"
    names[24 + 2 * n] = f"MATH half{n}: S/dP landed (WG{n & 1})"
    names[24 + 2 * n + 1] = f"MATH half{n}: P/dS published"
    names[56 + n] = f"PROD half{n}: stage free -> TMA issued"
for c in range(4):
    names[40 + 2 * c] = f"MATH item{c}: accumulators complete"
    names[40 + 2 * c + 1] = f"MATH item{c}: drained"
for ch in range(2):
    names[48 + 4 * ch] = f"MATHDETAIL half2 chunk{ch}: tmem ld issued + stats read"
    names[49 + 4 * ch] = f"MATHDETAIL half2 chunk{ch}: tmem ld landed"
    names[50 + 4 * ch] = f"MATHDETAIL half2 chunk{ch}: math done"
    names[51 + 4 * ch] = f"MATHDETAIL half2 chunk{ch}: smem stores done"
ev = sorted(((x, s) for s, x in enumerate(t) if x), key=lambda a: a[0])
t0 = ev[0][0]
for x, s in ev:
    print(f"{x - t0:8d} cyc  {names.get(s, s)}")
