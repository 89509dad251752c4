"""GEMM tile-rasterisation x L2-eviction-hint probe at the cfg-3 decoder shapes (round 2b).

    python tools/probe_gemm_l2.py                 # CUDA-event timings, one JSON line per (shape, raster, hints)
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 \
        --csv --log-file gpurun_out/l2probe_ncu.csv python tools/probe_gemm_l2.py --once     # one launch per case, same order

raster: -1 m-fastest, 0 automatic (pick_group_m), -2 bands for every multi-wave problem; hints: bit mask of
dalm_b200_gemm_set_l2_hints (1 A evict_last, 2 B evict_first, 4 stores evict_first). Operand sets rotate (3 x (A+B+out) >> L2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import _lib, ops

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32
once = "--once" in sys.argv
M = 18 * 256
L = 256
# (name, kind, N, K)   kind: rope | swiglu | f32res | bf16
SHAPES = [("qkv+rope", "rope", 12288, 4112), ("o_proj", "f32res", 4096, 4096), ("gate|up+swiglu", "swiglu", 22016, 4096),
          ("down", "f32res", 4096, 11008), ("lm_head", "bf16", 32000, 4096), ("dgrad_down", "bf16", 11008, 4096),
          ("dgrad_gu", "bf16", 4096, 22016), ("dgrad_o", "bf16", 4096, 4096), ("dgrad_qkv", "bf16", 4096, 12304)]
RASTERS = (-1, 0, -2)
HINTS = (0, 3, 7)
NBUF = 1 if once else 3


def aug(rows, cols):
    ld = (cols + 63) // 64 * 64 if cols % 64 else cols
    return (torch.randn(rows, ld, device=dev) * 0.1).to(bf16)[:, :cols]


def run_case(name, kind, N, K):
    As = [aug(M, K) for _ in range(NBUF)]
    Bs = [aug(N, K) for _ in range(NBUF)]
    out = torch.empty(M, N, device=dev, dtype=f32 if kind == "f32res" else bf16)
    out2 = torch.empty(M, N // 2, device=dev, dtype=bf16) if kind == "swiglu" else None
    res = torch.randn(M, N, device=dev) if kind == "f32res" else None
    cos = torch.rand(L, 64, device=dev); sin = torch.rand(L, 64, device=dev)
    i = [0]

    def fn():
        j = i[0] % NBUF; i[0] += 1
        if kind == "rope":
            ops.gemm_rope(As[j], Bs[j], cos, sin, L, 8192, out=out)
        elif kind == "swiglu":
            ops.gemm_swiglu(As[j], Bs[j], gu=out, act=out2)
        else:
            ops.gemm(As[j], Bs[j], out=out, resid=res)

    lib = _lib.load()
    for r in RASTERS:
        for h in HINTS:
            lib.dalm_b200_gemm_set_raster(r); lib.dalm_b200_gemm_set_l2_hints(h)
            if once:
                fn(); torch.cuda.synchronize()
                print(json.dumps({"case": name, "raster": r, "hints": h}), flush=True)
                continue
            for _ in range(3): fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(9): fn()
            e.record(); torch.cuda.synchronize()
            t = s.elapsed_time(e) / 9 * 1e-3
            print(json.dumps({"case": name, "M": M, "N": N, "K": K, "raster": r, "hints": h, "us": round(t * 1e6, 1),
                              "tflops": round(2.0 * M * N * K / t / 1e12, 1)}), flush=True)
    lib.dalm_b200_gemm_set_raster(0); lib.dalm_b200_gemm_set_l2_hints(-1)


if __name__ == "__main__":
    for sh in SHAPES:
        run_case(*sh)
