"""Micro-benchmarks of the hot kernels at cfg-3 shapes (CUDA events, L2 flushed between iterations by rotating buffers
larger than L2). Prints one line per kernel: time, TFLOP/s or GB/s."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def gemm_case(M, N, K, bn=0, out_dtype=bf16, resid=False, nbuf=4):
    As = [(torch.randn(M, K, device=dev) * 0.1).to(bf16) for _ in range(nbuf)]
    Bs = [(torch.randn(N, K, device=dev) * 0.1).to(bf16) for _ in range(nbuf)]
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    r = torch.randn(M, N, device=dev) if resid else None
    i = [0]
    def fn():
        j = i[0] % nbuf; i[0] += 1
        ops.gemm(As[j], Bs[j], out=out, resid=r, block_n=bn)
    t = timeit(fn)
    def ref():
        j = i[0] % nbuf; i[0] += 1
        torch.matmul(As[j], Bs[j].t())
    t_ref = timeit(ref)
    fl = 2.0 * M * N * K
    print(json.dumps({"kernel": "gemm", "M": M, "N": N, "K": K, "bn": bn, "ms": round(t * 1e3, 4), "tflops": round(fl / t / 1e12, 1),
                      "cublas_ms": round(t_ref * 1e3, 4), "cublas_tflops": round(fl / t_ref / 1e12, 1)}), flush=True)


if __name__ == "__main__":
    M = 18 * 256
    for (N, K) in [(12288, 4112), (4096, 4096), (22016, 4096), (4096, 11008), (32000, 4096), (4096, 12304), (4096, 22016), (11008, 4096), (4096, 32000)]:
        gemm_case(M, N, K, bn=256)
        gemm_case(M, N, K, bn=2256)
    gemm_case(M, 4096, 4096, bn=128)
    gemm_case(M, 4096, 4096, bn=2128)
    gemm_case(M, 4096, 4096, bn=256, out_dtype=f32, resid=True)
    for (Mb, N, K) in [(3204, 3072, 1048), (3204, 1024, 1024), (3204, 4096, 1024), (3204, 1024, 4096)]:
        gemm_case(Mb, N, K, bn=0)
        gemm_case(Mb, N, K, bn=64)
        gemm_case(Mb, N, K, bn=128)
        gemm_case(Mb, N, K, bn=2128)
    # attention fwd/bwd at cfg-3 decoder shape
    B, L, H, D = 18, 256, 32, 128
    qkv = torch.randn(B * L, 3 * H * D, device=dev).to(bf16)
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
    out, lse = ops.attention_fwd(q, k, v, mask, B, L, H, H, D, True)
    t = timeit(lambda: ops.attention_fwd(q, k, v, mask, B, L, H, H, D, True, out=out))
    fl = 4.0 * L * L * H * D * B
    print(json.dumps({"kernel": "attn_fwd_causal", "ms": round(t * 1e3, 4), "tflops_unhalved": round(fl / t / 1e12, 1)}), flush=True)
    do = torch.randn_like(out)
    dq = torch.empty_like(out); dk = torch.empty_like(out); dv = torch.empty_like(out)
    t = timeit(lambda: ops.attention_bwd(q, k, v, mask, out, lse, do, B, L, H, H, D, True, dq=dq, dk=dk, dv=dv))
    print(json.dumps({"kernel": "attn_bwd_causal", "ms": round(t * 1e3, 4), "tflops_unhalved": round(2 * fl / t / 1e12, 1)}), flush=True)
    # CE over vocab
    V = 32000
    logits = torch.randn(B, L, V, device=dev).to(bf16)
    ids = torch.randint(0, V, (B, L), device=dev)
    nsum = torch.tensor([float(B * (L - 1))], device=dev)
    dl = torch.empty_like(logits)
    from dalm_b200 import _lib
    def ce():
        _lib.call("dalm_b200_ce_marginal_fwd_bwd", logits.data_ptr(), dl.data_ptr(), 0, ids.data_ptr(), mask.data_ptr(), nsum.data_ptr(),
                  tok.data_ptr(), B, L, V, V, 1.0, torch.cuda.current_stream().cuda_stream)
    tok = torch.empty(B, L, device=dev)
    t = timeit(ce)
    by = 2.0 * B * (L - 1) * V * 2
    print(json.dumps({"kernel": "ce_rows", "ms": round(t * 1e3, 4), "gbs": round(by / t / 1e9, 1)}), flush=True)
    # in-batch loss
    qe = torch.nn.functional.normalize(torch.randn(18, 1024, device=dev), dim=1); pe = torch.nn.functional.normalize(torch.randn(18, 1024, device=dev), dim=1)
    qlen = torch.full((B,), 100, device=dev, dtype=torch.int64)
    cvec, ns = ops.marginal_counts(mask, qlen)
    t = timeit(lambda: ops.inbatch_loss(qe, pe, 100.0, cvec, ns), iters=50)
    print(json.dumps({"kernel": "inbatch_loss_B18", "us": round(t * 1e6, 2)}), flush=True)
    qe = torch.nn.functional.normalize(torch.randn(150, 1024, device=dev), dim=1); pe = torch.nn.functional.normalize(torch.randn(150, 1024, device=dev), dim=1)
    t = timeit(lambda: ops.inbatch_loss(qe, pe, 100.0), iters=50)
    print(json.dumps({"kernel": "inbatch_loss_B150", "us": round(t * 1e6, 2)}), flush=True)
