"""Quick GPU check of the pipelined backward kernels against the classic ones and fp64 (run under `timeout`: a pipeline
deadlock would otherwise spin forever). python tools/check_attn_pipe.py"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import _lib, ops

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
lib = _lib.load()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def ref64(q, k, v, mask, causal, B, L, Hq, Hkv, D, dm, d_out):
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    qh = qd.view(B, L, Hq, D).transpose(1, 2)
    kh = kd.view(B, L, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    vh = vd.view(B, L, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
    s = s.masked_fill(mask.view(B, 1, 1, L) == 0, float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(L, L, device=q.device, dtype=torch.bool), 1), float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    if dm is not None:
        p = p * dm
    o = (p @ vh).transpose(1, 2).reshape(B * L, Hq * D)
    o.backward(d_out.double())
    return qd.grad, kd.grad, vd.grad


cases = [(2, 256, 2, 2, 128, True, "none", 0.0), (3, 200, 4, 4, 128, True, "right", 0.0), (2, 96, 4, 4, 128, True, "left", 0.0),
         (1, 300, 4, 2, 128, True, "right", 0.0), (2, 384, 2, 2, 128, False, "right", 0.0), (18, 256, 32, 32, 128, True, "none", 0.0),
         (5, 50, 16, 16, 64, False, "right", 0.1), (4, 128, 16, 16, 64, False, "right", 0.1), (3, 37, 4, 4, 64, False, "none", 0.1),
         (2, 300, 7, 1, 64, True, "left", 0.0), (1, 1024, 71, 1, 64, True, "none", 0.0), (150, 50, 16, 16, 64, False, "right", 0.1),
         (2, 200, 2, 2, 64, False, "right", 0.1), (2, 384, 4, 2, 64, True, "right", 0.0), (3, 100, 2, 2, 128, False, "left", 0.0)]
worst = 0.0
for (B, L, Hq, Hkv, D, causal, pad, pd) in cases:
    torch.manual_seed(B * 1000 + L + Hq)
    wide = (Hq + 2 * Hkv) * D
    qkv = torch.randn(B * L, wide, device=dev).to(bf16)
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    if pad == "right":
        for b in range(B): mask[b, L - 3 - (5 * b) % (L // 2):] = 0
    elif pad == "left":
        for b in range(B): mask[b, :4 + 3 * b] = 0
    d = ops.Drop(pd, 77, (5 << 8) | 9, None) if pd > 0 else None
    out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, Hq, Hkv, D, causal, drop=d)
    rows = mask.bool().view(-1) if (causal and pad == "left") else torch.ones(B * L, dtype=torch.bool, device=dev)
    d_out = torch.randn(B * L, Hq * D, device=dev).to(bf16)
    d_out[~rows] = 0
    res = {}
    for mode in (1, 0):
        lib.dalm_b200_attention_tc_set_mode(mode)
        t0 = time.time()
        dq, dk, dv = ops.attention_tc_bwd(q, k, v, mask, out, lse, d_out, B, L, Hq, Hkv, D, causal, drop=d)
        torch.cuda.synchronize()
        res[mode] = (dq.float(), dk.float(), dv.float(), time.time() - t0)
    lib.dalm_b200_attention_tc_set_mode(1)
    line = f"B{B} L{L} H{Hq}/{Hkv} D{D} causal={causal} pad={pad} p={pd}: pipe-vs-classic " + \
        " ".join(f"{rel(res[1][i], res[0][i]):.1e}" for i in range(3))
    if B * Hq * L * L <= 4e8:
        dm = None
        if d is not None:
            Lp = (L + 7) // 8 * 8
            dm = ops.dropout_scale(B * Hq * L * Lp, d, dev).view(B, Hq, L, Lp)[..., :L].double()
        gq, gk, gv = ref64(q, k, v, mask, causal, B, L, Hq, Hkv, D, dm, d_out)
        e = [rel(res[1][0], gq), rel(res[1][1], gk), rel(res[1][2], gv)]
        worst = max(worst, *e)
        line += "  vs-fp64 " + " ".join(f"{x:.1e}" for x in e)
    print(line, flush=True)
print("WORST_VS_FP64", worst, "OK" if worst < 4e-2 else "FAIL", flush=True)


def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return round(s.elapsed_time(e) / n * 1e3, 1)


for tag, (B, L, Hq, Hkv, D, causal, pd) in {"llama": (18, 256, 32, 32, 128, True, 0.0), "bge_q": (150, 50, 16, 16, 64, False, 0.1),
                                             "bge_p": (150, 128, 16, 16, 64, False, 0.1), "bge_q18": (18, 50, 16, 16, 64, False, 0.1),
                                             "bge_p18": (18, 128, 16, 16, 64, False, 0.1), "falcon": (4, 2048, 71, 1, 64, True, 0.0)}.items():
    qkv = torch.randn(B * L, (Hq + 2 * Hkv) * D, device=dev).to(bf16)
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    d = ops.Drop(pd, 1, 2, None) if pd > 0 else None
    out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, Hq, Hkv, D, causal, drop=d)
    do = torch.randn_like(out); dq = torch.empty_like(q); dk = torch.empty_like(k); dv = torch.empty_like(v)
    r = {}
    for mode in (1, 0):
        lib.dalm_b200_attention_tc_set_mode(mode)
        r["pipe_us" if mode else "classic_us"] = t(lambda: ops.attention_tc_bwd(q, k, v, mask, out, lse, do, B, L, Hq, Hkv, D, causal, dq=dq, dk=dk, dv=dv, drop=d))
    lib.dalm_b200_attention_tc_set_mode(1)
    print("BWD_TIMING", tag, r, flush=True)
