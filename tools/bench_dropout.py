"""per-kernel cost of the dropout sites at cfg-3 shapes (CUDA events): each kernel with and without its DropCfg"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops
from tools.bench_kernels import timeit
dev = torch.device("cuda:0"); bf16 = torch.bfloat16; f32 = torch.float32
off = torch.zeros(1, dtype=torch.int64, device=dev)
D = lambda p, s: ops.Drop(p, 7, s, off)

def rep(name, fn0, fn1, iters=30):
    t0, t1 = timeit(fn0, iters=iters), timeit(fn1, iters=iters)
    print(json.dumps({"kernel": name, "us_nodrop": round(t0 * 1e6, 1), "us_drop": round(t1 * 1e6, 1)}), flush=True)

# Llama LoRA sites: M=4608, H=4096
M, H, R = 4608, 4096, 16
buf = torch.zeros(M, H + 64, device=dev, dtype=bf16); buf[:, :H] = torch.randn(M, H, device=dev).to(bf16)
x = buf[:, :H]; a_stack = torch.randn(64, H, device=dev).to(bf16)
rep("skinny_gemm llama", lambda: ops.skinny_gemm(x, a_stack, buf[:, H:], K=H, R=R), lambda: ops.skinny_gemm(x, a_stack, buf[:, H:], K=H, R=R, dropx=D(0.05, 3)))
g = torch.randn(M, 16, device=dev).to(bf16); o0 = torch.zeros(8, H, device=dev); o1 = torch.zeros(8, H, device=dev)
rep("lora_wgrad dA llama", lambda: ops.lora_wgrad_(x, g, o0, H, 1, H, 16, 1.0, out1=o1), lambda: ops.lora_wgrad_(x, g, o0, H, 1, H, 16, 1.0, out1=o1, dropx=D(0.05, 3)))
dh = torch.randn(M, H, device=dev).to(bf16)
t = timeit(lambda: ops.lora_dx_(dh, g, a_stack, K=H, R=16, drop=D(0.05, 3)), iters=30)
print(json.dumps({"kernel": "lora_dx llama", "us_drop": round(t * 1e6, 1)}), flush=True)
dq = torch.zeros(M, 12288 + 64, device=dev, dtype=bf16)[:, :12288 + 16]; wt = torch.zeros(H, 12288 + 64, device=dev, dtype=bf16)[:, :12288 + 16]
rep("dgrad gemm folded vs unfolded", lambda: ops.gemm(dq, wt), lambda: ops.gemm(dq[:, :12288], wt[:, :12288]))
# BERT sites: M=3204, H=1024
Mb, Hb = 3204, 1024
z = torch.randn(Mb, Hb, device=dev); gm = torch.randn(Hb, device=dev); be = torch.randn(Hb, device=dev)
rep("layernorm_fwd bert", lambda: ops.layernorm_fwd(z, gm, be, 1e-12), lambda: ops.layernorm_fwd(z, gm, be, 1e-12, drop=D(0.1, 0)))
y32, y16, mean, rstd = ops.layernorm_fwd(z, gm, be, 1e-12); dy = torch.randn(Mb, Hb, device=dev)
rep("layernorm_bwd bert", lambda: ops.layernorm_bwd(z, gm, mean, rstd, dy_f32=dy), lambda: ops.layernorm_bwd(z, gm, mean, rstd, dy_f32=dy, drop16=D(0.1, 1)))
a = torch.randn(Mb, Hb, device=dev).to(bf16); w = torch.randn(Hb, Hb, device=dev).to(bf16); bias = torch.randn(Hb, device=dev); res = torch.randn(Mb, Hb, device=dev)
out = torch.empty(Mb, Hb, device=dev)
rep("gemm bert self-output", lambda: ops.gemm(a, w, out=out, bias=bias, resid=res), lambda: ops.gemm(a, w, out=out, bias=bias, resid=res, drop=D(0.1, 1)))
for (B, L) in ((18, 128), (18, 50)):
    nh, hd = 16, 64
    qkv = torch.randn(B * L, 3 * Hb, device=dev).to(bf16); mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    q, k, v = qkv[:, :Hb], qkv[:, Hb:2 * Hb], qkv[:, 2 * Hb:]
    o, lse = ops.attention_fwd(q, k, v, mask, B, L, nh, nh, hd, False)
    rep(f"attn_fwd bert L={L}", lambda: ops.attention_fwd(q, k, v, mask, B, L, nh, nh, hd, False, out=o), lambda: ops.attention_fwd(q, k, v, mask, B, L, nh, nh, hd, False, out=o, drop=D(0.1, 8)))
    do = torch.randn_like(o); dqq = torch.empty_like(o); dk = torch.empty_like(o); dv = torch.empty_like(o)
    rep(f"attn_bwd bert L={L}", lambda: ops.attention_bwd(q, k, v, mask, o, lse, do, B, L, nh, nh, hd, False, dq=dqq, dk=dk, dv=dv),
        lambda: ops.attention_bwd(q, k, v, mask, o, lse, do, B, L, nh, nh, hd, False, dq=dqq, dk=dk, dv=dv, drop=D(0.1, 8)))
