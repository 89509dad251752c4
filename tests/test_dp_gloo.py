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
