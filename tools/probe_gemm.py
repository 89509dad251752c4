import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops
from tools.bench_kernels import timeit
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
M = 4608
for (N, K) in [(4096, 4096), (4096, 32000), (22016, 4096), (12288, 4112), (32000, 4096), (4096, 11008)]:
    a = (torch.randn(M, K, device=dev) * 0.1).to(bf16); b = (torch.randn(N, K, device=dev) * 0.1).to(bf16)
    out = torch.empty(M, N, device=dev, dtype=bf16)
    for bn in (256, 2256):
        for act in (0,):
            t = timeit(lambda: ops.gemm(a, b, out=out, block_n=bn, act=act), iters=30)
            print(json.dumps({"N": N, "K": K, "bn": bn, "act": act, "ms": round(t * 1e3, 4), "tflops": round(2.0 * M * N * K / t / 1e12, 1)}), flush=True)
    o32 = torch.empty(M, N, device=dev, dtype=torch.float32); r32 = torch.randn(M, N, device=dev)
    t = timeit(lambda: ops.gemm(a, b, out=o32, resid=r32, block_n=256), iters=30)
    print(json.dumps({"N": N, "K": K, "bn": 256, "epilogue": "f32+resid", "tflops": round(2.0 * M * N * K / t / 1e12, 1)}), flush=True)
    t = timeit(lambda: torch.matmul(a, b.t()), iters=30)
    print(json.dumps({"N": N, "K": K, "cublas_tflops": round(2.0 * M * N * K / t / 1e12, 1)}), flush=True)
