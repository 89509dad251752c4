"""Tile-width probe for the small encoder GEMMs (M = 3204 token rows of cfg-3): block_n 64 / 128 / 256 / auto, CUDA events,
L2 flushed between launches by cycling through operand copies larger than L2.  python tools/probe_tiles.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops

dev = torch.device("cuda:0")
shapes = [(3204, 1024, 4096), (3204, 1024, 1024), (3204, 1024, 3072), (3204, 3072, 1048), (3204, 4096, 1024), (900, 1024, 1024),
          (26700, 1024, 4096)]
for M, N, K in shapes:
    ncopy = max(2, int(300e6 // ((M * K + N * K) * 2)) + 1)
    A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(ncopy)]
    B = [torch.randn(N, K, device=dev).bfloat16() for _ in range(ncopy)]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    row = {"M": M, "N": N, "K": K}
    for bn in (0, 64, 128, 256):
        for i in range(3):
            ops.gemm(A[i % ncopy], B[i % ncopy], out=out, block_n=bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for i in range(reps):
            ops.gemm(A[i % ncopy], B[i % ncopy], out=out, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        row[f"bn{bn}_us"] = round(us, 1)
        row[f"bn{bn}_tflops"] = round(2.0 * M * N * K / us / 1e6, 1)
    print(json.dumps(row), flush=True)
