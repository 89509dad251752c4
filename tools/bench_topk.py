"""Roofline of the evaluation search kernel (csrc/topk.cu): exact inner-product top-k of a query batch over the resident
passage embeddings at the reference's evaluation sizes (200k unique passages x 1024, test_batch_size 8, top_k 10;
dalm/eval/eval_retriever_only.py:76-97). HBM-bound: algorithmic bytes = N*D*4 per query tile of 8.
    python tools/bench_topk.py [N] [D] [nq] [K]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalm_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 8
K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
P = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=1)          # 819 MB > 126 MB L2
Q = torch.nn.functional.normalize(torch.randn(nq, D, device=dev, generator=g), dim=1)
for _ in range(3):
    ops.topk_ip(Q, P, K)
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    s, i = ops.topk_ip(Q, P, K)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
tiles = (nq + 7) // 8
algo = N * D * 4 * tiles
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = (peak or {}).get("hbm_gbs_sustained") or (peak or {}).get("hbm_gbs") or 6650.0
ref = (Q @ P.t()).topk(K, dim=1)
print(json.dumps({"kernel": "topk_scan_kernel + topk_merge_kernel", "N": N, "D": D, "nq": nq, "K": K, "ms": ms,
                  "queries_per_s": nq / (ms * 1e-3), "algorithmic_bytes": algo, "achieved_gbs": algo / (ms * 1e-3) / 1e9,
                  "peak_gbs": hbm, "frac": algo / (ms * 1e-3) / 1e9 / hbm,
                  "indices_equal_torch_topk": bool(torch.equal(ref.indices.int(), i))}))
