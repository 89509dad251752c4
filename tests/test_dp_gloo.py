"""CPU, world_size 2 on gloo: the data-parallel plumbing of the trainers (gradient averaging == DDP mean, scalar loss
sum, disjoint rank-strided batches). The reference gets these from accelerate/DDP implicitly
(train_rage2e.py:416-418,469,471)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dalm_b200.accel import Accelerator
    acc = Accelerator(cpu=True)
    assert acc.num_processes == world and acc.process_index == rank
    # gradient averaging of two flat buffers == mean over ranks (what DDP does)
    g1 = torch.full((1000,), float(rank + 1)); g2 = torch.arange(10.0) * (rank + 1)
    acc.average_gradients([g1, g2])
    ok = torch.allclose(g1, torch.full((1000,), 1.5)) and torch.allclose(g2, torch.arange(10.0) * 1.5)
    # not averaged while accumulating
    acc.sync_gradients = False
    g3 = torch.full((4,), float(rank))
    acc.average_gradients([g3])
    ok = ok and torch.equal(g3, torch.full((4,), float(rank)))
    acc.sync_gradients = True
    # scalar loss: rank SUM (reference :469)
    tot = acc.reduce(torch.tensor(float(rank + 1)), reduction="sum").item()
    ok = ok and tot == 3.0
    # loader sharding: same permutation on both ranks, disjoint batches
    ds = list(range(23))
    gen = torch.Generator().manual_seed(42)
    dl = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=True, generator=gen)
    sl = acc.prepare(dl)
    mine = [b.tolist() for b in sl]
    gathered = [None, None]
    dist.all_gather_object(gathered, mine)
    flat0, flat1 = sum(gathered[0], []), sum(gathered[1], [])
    ok = ok and len(gathered[0]) == len(gathered[1]) == 3 and not (set(flat0[:8]) & set(flat1[:8]))
    ok = ok and set(flat0) | set(flat1) == set(range(23))
    # the step's single-collective reducer: both LoRA banks' gradients + the loss scalar in ONE all-reduce
    from dalm_b200.engine.lora import LoraBank
    b1, b2 = LoraBank([("m.q", 16, 24)], device="cpu", seed=1), LoraBank([("m.v", 16, 8), ("m.k", 16, 8)], device="cpu", seed=2)
    prm = torch.nn.Parameter(b1.flat); prm.grad = b1.grad; b1.param = prm
    sync = acc.gradient_sync([b1, b2])
    assert prm.grad.data_ptr() == b1.grad.data_ptr() == sync.arena.data_ptr()        # re-homed into the arena
    b1.gA["m.q"].fill_(float(rank + 1)); b1.gB["m.q"].fill_(10.0 * (rank + 1))      # written through the per-adapter views
    b2.gB["m.k"].fill_(100.0 * (rank + 1))
    n0 = sync.collectives
    tot = sync.reduce(torch.tensor(2.0 * (rank + 1)))
    ok = ok and sync.collectives == n0 + 1 and abs(tot.item() - 6.0) < 1e-6          # loss: rank SUM
    ok = ok and torch.allclose(b1.gA["m.q"], torch.full((8, 16), 1.5)) and torch.allclose(b1.gB["m.q"], torch.full((24, 8), 15.0))
    ok = ok and torch.allclose(b2.gB["m.k"], torch.full((8, 8), 150.0)) and b2.gA["m.v"].abs().max().item() == 0.0
    b1.zero_grad()
    ok = ok and sync.arena[: b1.grad.numel()].abs().max().item() == 0.0
    # full fine-tuning: per-layer buckets announced during the backward are averaged right away (DDP bucket hooks), the
    # rest (embeddings / norms) at the end of the step; every element is averaged exactly once
    from dalm_b200.engine.dense import DenseBank
    dbank = DenseBank([("embed", (10, 8), "acc"), ("L0.Wa", (4, 8), "gemm"), ("L0.Wb", (8, 8), "gemm"), ("L1.Wa", (4, 8), "gemm"),
                       ("L1.Wb", (8, 8), "gemm"), ("lm_head", (10, 8), "gemm"), ("L0.g", (8,), "acc")], "cpu")
    sync2 = acc.gradient_sync([dbank])
    ok = ok and sync2.overlaps_backward and not acc.gradient_sync([b1]).overlaps_backward
    dbank.g32.fill_(float(rank + 1))
    c0 = sync2.collectives
    dbank.bucket_ready("lm_head"); dbank.bucket_ready("L1."); dbank.bucket_ready("L0.")
    ok = ok and sync2.collectives == c0 + 3 and torch.allclose(dbank.g("L1.Wb"), torch.full((8, 8), 1.5))
    ok = ok and torch.allclose(dbank.g("embed"), torch.full((10, 8), float(rank + 1)))      # not announced: still local
    sync2.reduce(torch.tensor(1.0))
    ok = ok and torch.allclose(dbank.g32, torch.full_like(dbank.g32, 1.5)) and dbank.reduced == []
    sync2.armed = False                                                                        # accumulation micro-step
    dbank.g32.fill_(float(rank + 1)); c1 = sync2.collectives
    dbank.bucket_ready("L0.")
    ok = ok and sync2.collectives == c1 and torch.allclose(dbank.g("L0.Wa"), torch.full((4, 8), float(rank + 1)))
    acc.wait_for_everyone()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_data_parallel_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------------------------------------------------------------
# cross-rank in-batch negatives (optional extension): the host logic on gloo with the fused kernel replaced by its fp64
# restatement - every rank's slice of the global gradient, DDP-averaged, must equal the gradient of the rank-mean objective
# ---------------------------------------------------------------------------------------------------------------------
def _inbatch_oracle(q, p, scale, cvec, nsum, need_grad=True, grad_out=1.0):
    """dalm_b200.ops.inbatch_loss restated with torch autograd in fp64 (same outputs, same argument meaning)"""
    q64, p64 = q.double().requires_grad_(True), p.double().requires_grad_(True)
    S = (q64 @ p64.t()) * scale
    n = S.shape[0]
    ar = torch.arange(n)
    rows = torch.log_softmax(S, 1)[ar, ar]
    cols = torch.log_softmax(S, 0)[ar, ar]
    lc = -(rows + cols).mean() / 2.0
    doc = -(cvec.double() * rows).sum() / nsum.double()[0] if cvec is not None else torch.zeros((), dtype=torch.float64)
    ((lc + doc) * grad_out).backward()
    return {"S": S.detach().float(), "dlp": rows.detach().float(),
            "losses": torch.stack([lc, doc, lc + doc, nsum.double()[0] if nsum is not None else torch.zeros((), dtype=torch.float64)]).detach().float(),
            "dQ": q64.grad.float(), "dP": p64.grad.float()}


def _neg_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      DALM_B200_CROSS_RANK_NEGATIVES="1")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dalm_b200.accel import Accelerator
    from dalm_b200.training.utils import negatives
    Accelerator(cpu=True)
    assert negatives.active()
    D, scale = 16, 20.0
    Bs = [5, 3]                                                      # a short final batch on rank 1
    g = torch.Generator().manual_seed(7)
    Q = [torch.nn.functional.normalize(torch.randn(b, D, generator=g), dim=1) for b in Bs]
    P = [torch.nn.functional.normalize(torch.randn(b, D, generator=g), dim=1) for b in Bs]
    C = [torch.rand(b, generator=g) * 10 for b in Bs]
    N = [torch.tensor([37.0]), torch.tensor([21.0])]
    ok = True
    for marg in (True, False):
        r = negatives.global_inbatch_loss(Q[rank], P[rank], scale, C[rank] if marg else None, N[rank] if marg else None, True, 0.5,
                                          rank=rank, world=world, loss_fn=_inbatch_oracle)
        # the objective every rank's gradients must add up to: J = (1/W) sum_r [Lc(S_global) + doc_r]
        qa = torch.cat(Q).double().requires_grad_(True); pa = torch.cat(P).double().requires_grad_(True)
        S = (qa @ pa.t()) * scale
        n = S.shape[0]; ar = torch.arange(n)
        rows = torch.log_softmax(S, 1)[ar, ar]; cols = torch.log_softmax(S, 0)[ar, ar]
        lc = -(rows + cols).mean() / 2.0
        J = lc
        docs = []
        if marg:
            off = 0
            for b, c, nn in zip(Bs, C, N):
                docs.append(-(c.double() * rows[off:off + b]).sum() / nn.double()[0]); off += b
            J = J + sum(docs) / world
        (J * 0.5).backward()
        lo = sum(Bs[:rank])
        # what a data-parallel MEAN over ranks of "this rank's rows only" yields == dJ/dq: each rank holds W x its slice
        ok = ok and torch.allclose(r["dQ"].double() / world, qa.grad[lo:lo + Bs[rank]], atol=1e-6)
        ok = ok and torch.allclose(r["dP"].double() / world, pa.grad[lo:lo + Bs[rank]], atol=1e-6)
        ok = ok and abs(r["losses"][0].item() - lc.item()) < 1e-5 and r["S"].shape == (8, 8)
        if marg:
            ok = ok and abs(r["losses"][1].item() - docs[rank].item()) < 1e-5 and abs(r["losses"][3].item() - N[rank].item()) < 1e-6
            ok = ok and torch.allclose(r["dlp"].double(), rows[lo:lo + Bs[rank]].detach(), atol=1e-5)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_cross_rank_negatives_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_neg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
