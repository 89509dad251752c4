import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops, _lib
B, L, H, D = 18, 256, 32, 128
dev = torch.device("cuda:0")
qkv = torch.randn(B * L, 3 * H * D, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
mask = torch.ones(B, L, dtype=torch.int64, device=dev)
for _ in range(3): ops.attention_tc_fwd(q, k, v, mask, B, L, H, H, D, True)
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
_lib.load().dalm_b200_attention_tc_set_debug(dbg.data_ptr())
ops.attention_tc_fwd(q, k, v, mask, B, L, H, H, D, True)
torch.cuda.synchronize()
_lib.load().dalm_b200_attention_tc_set_debug(None)
t = dbg.cpu().tolist(); t0 = t[0]
names = {0: "start", 1: "setup done", 40: "before final sync", 41: "end"}
for j in range(2):
    names.update({2 + j*8: f"ctl j{j} loads issued", 3 + j*8: f"ctl j{j} loads landed", 4 + j*8: f"ctl j{j} S issued", 5 + j*8: f"ctl j{j} P ready",
                  6 + j*8: f"ctl j{j} PV issued", 20 + j*8: f"cmp j{j} S landed", 21 + j*8: f"cmp j{j} pass1 done", 22 + j*8: f"cmp j{j} P published",
                  23 + j*8: f"cmp j{j} PV landed", 24 + j*8: f"cmp j{j} O accumulated"})
for slot, ts in sorted(((s, x) for s, x in enumerate(t) if x), key=lambda a: a[1]):
    print(f"{ts - t0:8d} cyc  {names.get(slot, slot)}")
