"""-m gpu: the fused loss kernels against the oracle (oracle/losses.py: CPU fp64 restatement of reference
dalm/training/utils/train_utils.py:76-138, itself pinned to the reference's outputs by tests/test_oracle_golden.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(B, D, L, V, seed, pad="right", dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1)
    p = torch.nn.functional.normalize(torch.randn(B, D, generator=g) + 0.5 * q, dim=1)
    logits = (torch.randn(B, L, V, generator=g) * 2).to(dtype)
    ids = torch.randint(0, V, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    for b in range(B):
        n_pad = int(torch.randint(0, max(1, L // 3), (1,), generator=g))
        if n_pad:
            if pad == "right": mask[b, L - n_pad:] = 0
            elif pad == "left": mask[b, :n_pad] = 0
    qlen = torch.randint(1, L + 4, (B,), generator=g)
    qlen[0] = 1
    if B > 1: qlen[1] = L - 1
    if B > 2: qlen[2] = L
    if B > 3: qlen[3] = L + 3
    return q, p, logits, ids, mask, qlen


@pytest.mark.parametrize("B,D,L,V,pad", [(2, 384, 16, 1000, "right"), (18, 1024, 64, 32000, "right"),
                                         (18, 1024, 32, 32000, "left"), (5, 64, 9, 131, "right"), (150, 1024, 8, 512, "right")])
def test_rag_loss_matches_oracle(cuda_dev, B, D, L, V, pad):
    from dalm_b200 import ops
    from oracle import losses
    q, p, logits, ids, mask, qlen = _case(B, D, L, V, seed=B * 7 + L, pad=pad)
    ref = losses.rag_loss_and_grads(q, p, 100.0, logits.float(), ids, mask, qlen)
    dev = cuda_dev
    qd, pd, lg, idd, md, qld = (t.to(dev) for t in (q, p, logits, ids, mask, qlen))
    cvec, nsum = ops.marginal_counts(md, qld)
    c_ref, n_ref = losses.marginal_counts(mask, qlen)
    assert torch.equal(cvec.cpu().double(), c_ref) and nsum.item() == n_ref.item()     # integer counts: exact
    r = ops.inbatch_loss(qd, pd, 100.0, cvec, nsum)
    tok_lp, dl = ops.ce_marginal(lg, idd, md, nsum)
    out = ops.finalize_loss(tok_lp, md, nsum, r["losses"])
    # fp32 kernels vs fp64 oracle
    assert torch.allclose(r["S"].cpu().double(), ref["S"], rtol=1e-5, atol=1e-4)
    assert abs(out[0].item() - ref["Lc"].item()) < 1e-4 * max(1, abs(ref["Lc"].item()))
    assert abs(out[1].item() - ref["Lm"].item()) < 1e-4 * max(1, abs(ref["Lm"].item()))
    assert abs(out[2].item() - ref["loss"].item()) < 1e-4 * max(1, abs(ref["loss"].item()))
    gq, gp = r["dQ"].cpu().double(), r["dP"].cpu().double()
    assert (gq - ref["dQ"]).norm() / ref["dQ"].norm() < 1e-4
    assert (gp - ref["dP"]).norm() / ref["dP"].norm() < 1e-4
    # dlogits are emitted in bf16: 2^-8 relative
    gl = dl.float().cpu().double()
    assert (gl - ref["dlogits"]).norm() / ref["dlogits"].norm() < 6e-3


def test_retriever_only_loss(cuda_dev):
    from dalm_b200 import ops
    from oracle import losses
    q, p, *_ = _case(150, 1024, 4, 8, seed=3)
    ref = losses.retriever_loss_and_grads(q, p, 100.0)
    r = ops.inbatch_loss(q.to(cuda_dev), p.to(cuda_dev), 100.0)
    assert abs(r["losses"][0].item() - ref["loss"].item()) < 1e-4 * abs(ref["loss"].item())
    assert r["losses"][1].item() == 0.0
    assert (r["dQ"].cpu().double() - ref["dQ"]).norm() / ref["dQ"].norm() < 1e-4
    assert (r["dP"].cpu().double() - ref["dP"]).norm() / ref["dP"].norm() < 1e-4


def test_ce_fp32_inplace_and_large_vocab(cuda_dev):
    """fp32 logits (the reference's dtype without mixed precision), in-place gradient, and a vocabulary too large for
    the shared-memory row cache (Falcon's 65024 in fp32 = 254 KB)"""
    from dalm_b200 import ops
    from oracle import losses
    q, p, logits, ids, mask, qlen = _case(3, 64, 6, 65024, seed=11, dtype=torch.float32)
    ref = losses.rag_loss_and_grads(q, p, 100.0, logits, ids, mask, qlen)
    dev = cuda_dev
    cvec, nsum = ops.marginal_counts(mask.to(dev), qlen.to(dev))
    r = ops.inbatch_loss(q.to(dev), p.to(dev), 100.0, cvec, nsum)
    lg = logits.to(dev).clone()
    tok_lp, dl = ops.ce_marginal(lg, ids.to(dev), mask.to(dev), nsum, inplace=True)
    assert dl.data_ptr() == lg.data_ptr()
    out = ops.finalize_loss(tok_lp, mask.to(dev), nsum, r["losses"])
    assert abs(out[2].item() - ref["loss"].item()) < 1e-5 * abs(ref["loss"].item())
    assert (dl.cpu().double() - ref["dlogits"]).norm() / ref["dlogits"].norm() < 1e-5


def test_cross_rank_negatives_on_the_fused_kernel(cuda_dev):
    """DALM_B200_CROSS_RANK_NEGATIVES: two emulated ranks (gather injected) through the real fused in-batch kernel; each rank's
    slice of dQ / dP, divided by the world size (the data-parallel mean), equals the fp64 gradient of
    J = (1/W) sum_r [Lc(S_global) + doc_r] for its rows; the reported losses are (global Lc, this rank's doc term)"""
    from dalm_b200 import ops
    from dalm_b200.training.utils import negatives
    dev = cuda_dev
    W, Bs, D, scale, gout = 2, [18, 18], 1024, 100.0, 1.0
    g = torch.Generator().manual_seed(3)
    Q = [torch.nn.functional.normalize(torch.randn(b, D, generator=g), dim=1) for b in Bs]
    P = [torch.nn.functional.normalize(torch.randn(b, D, generator=g) + 0.5 * q, dim=1) for b, q in zip(Bs, Q)]
    C = [torch.randint(0, 60, (b,), generator=g).float() for b in Bs]
    N = [torch.tensor([700.0]), torch.tensor([655.0])]
    qa = torch.cat(Q).double().requires_grad_(True); pa = torch.cat(P).double().requires_grad_(True)
    S = (qa @ pa.t()) * scale
    ar = torch.arange(S.shape[0])
    rows = torch.log_softmax(S, 1)[ar, ar]; cols = torch.log_softmax(S, 0)[ar, ar]
    lc = -(rows + cols).mean() / 2.0
    docs, off = [], 0
    for b, c, n in zip(Bs, C, N):
        docs.append(-(c.double() * rows[off:off + b]).sum() / n.double()[0]); off += b
    (lc + sum(docs) / W).backward()
    for rank in range(W):
        def gather(t, rank=rank):
            # what the all-gather returns on this rank: every rank's block of the tensor being gathered, identified by its shape
            for src in (Q, P):
                if t.dim() == 2 and torch.equal(t.cpu(), src[rank]):
                    return [x.to(dev) for x in src]
            return [(c / n).to(dev) for c, n in zip(C, N)]
        r = negatives.global_inbatch_loss(Q[rank].to(dev), P[rank].to(dev), scale, C[rank].to(dev), N[rank].to(dev), True, gout,
                                          rank=rank, world=W, loss_fn=ops.inbatch_loss, gather=gather)
        lo = sum(Bs[:rank])
        rel = lambda a, b: ((a.double().cpu() - b).norm() / (b.norm() + 1e-30)).item()
        assert rel(r["dQ"] / W, qa.grad[lo:lo + Bs[rank]]) < 1e-4 and rel(r["dP"] / W, pa.grad[lo:lo + Bs[rank]]) < 1e-4
        assert abs(r["losses"][0].item() - lc.item()) < 1e-4 * abs(lc.item()) + 1e-6
        assert abs(r["losses"][1].item() - docs[rank].item()) < 1e-4 * abs(docs[rank].item()) + 1e-6
        assert r["S"].shape == (36, 36) and r["dlp"].shape == (18,)
