"""GEMM tuning probe (round 2): tile rasterisation bands x epilogue kinds at the cfg-3 generator shapes, L2 defeated by
rotating operand sets (4 x (A + B) >> 126 MB). CUDA events, one JSON line per case. Usage: python tools/bench_gemm_r2.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import _lib, ops

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32


def timeit(fn, iters=12, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def case(M, N, K, kind, group_m, nbuf=4, ref=False):
    As = [(torch.randn(M, K, device=dev) * 0.1).to(bf16) for _ in range(nbuf)]
    Bs = [(torch.randn(N, K, device=dev) * 0.1).to(bf16) for _ in range(nbuf)]
    out = torch.empty(M, N, device=dev, dtype=f32 if kind != "bf16" else bf16)
    r = torch.randn(M, N, device=dev) if kind == "f32+resid" else None
    _lib.load().dalm_b200_gemm_set_raster(group_m)
    i = [0]
    def fn():
        j = i[0] % nbuf; i[0] += 1
        ops.gemm(As[j], Bs[j], out=out, resid=r)
    t = timeit(fn)
    _lib.load().dalm_b200_gemm_set_raster(0)
    row = {"M": M, "N": N, "K": K, "kind": kind, "group_m": group_m, "us": round(t * 1e6, 1), "tflops": round(2.0 * M * N * K / t / 1e12, 1)}
    if ref:
        def rf():
            j = i[0] % nbuf; i[0] += 1
            torch.matmul(As[j], Bs[j].t())
        tr = timeit(rf)
        row["cublas_tflops"] = round(2.0 * M * N * K / tr / 1e12, 1)
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    M = 18 * 256
    shapes = [(22016, 4096, "bf16"), (4096, 11008, "f32+resid"), (12288, 4112, "bf16"), (4096, 4096, "f32+resid"),
              (4096, 4096, "bf16"), (4096, 4096, "f32"), (11008, 4096, "bf16"), (4096, 22016, "bf16"), (4096, 12304, "bf16"), (32000, 4096, "bf16")]
    for N, K, kind in shapes:
        for gm in (-1, 0, 9, 12, 18):
            case(M, N, K, kind, gm, ref=(gm == -1))
    for (Mb, N, K, kind) in [(3204, 3072, 1048, "bf16"), (3204, 1024, 1024, "f32+resid"), (3204, 4096, 1024, "bf16"), (3204, 1024, 4096, "f32+resid"),
                             (26700, 3072, 1048, "bf16"), (26700, 1024, 4096, "f32+resid")]:
        for gm in (-1, 0):
            case(Mb, N, K, kind, gm, ref=(gm == -1))
