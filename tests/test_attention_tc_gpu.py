"""-m gpu: tcgen05/TMEM attention (head_dim 128) against the fp64 torch reference and the mma.sync kernels."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _ref(q, k, v, mask, causal, B, L, Hq, Hkv, D):
    qh = q.double().view(B, L, Hq, D).transpose(1, 2)
    kh = k.double().view(B, L, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    vh = v.double().view(B, L, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
    if mask is not None:
        s = s.masked_fill(mask.view(B, 1, 1, L) == 0, float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(L, L, device=q.device, dtype=torch.bool), 1), float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    return (p @ vh).transpose(1, 2).reshape(B * L, Hq * D)


@pytest.mark.parametrize("B,L,Hq,Hkv,causal,pad", [
    (2, 256, 2, 2, True, "none"), (2, 128, 2, 2, True, "none"), (3, 200, 4, 4, True, "right"), (2, 96, 4, 4, True, "left"),
    (1, 300, 4, 2, True, "right"), (2, 384, 2, 2, False, "right"), (18, 256, 32, 32, True, "none"),
])
def test_attention_tc_forward(cuda_dev, B, L, Hq, Hkv, causal, pad):
    from dalm_b200 import ops
    D = 128
    torch.manual_seed(B * 100 + L)
    dev = cuda_dev
    qkv = torch.randn(B * L, (Hq + 2 * Hkv) * D, device=dev).to(bf16)
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    if pad == "right":
        for b in range(B): mask[b, L - 3 - 5 * b:] = 0
    elif pad == "left":
        for b in range(B): mask[b, :4 + 3 * b] = 0
    out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, Hq, Hkv, D, causal)
    out2, lse2 = ops.attention_fwd(q, k, v, mask, B, L, Hq, Hkv, D, causal)
    if B * L * Hq <= 20000:
        ref = _ref(q, k, v, mask, causal, B, L, Hq, Hkv, D)
        rows = mask.bool().view(-1) if (causal and pad == "left") else torch.ones(B * L, dtype=torch.bool, device=dev)
        assert _rel(out.float()[rows], ref[rows]) < 1.5e-2
        if causal and pad == "left":
            assert out.float()[~rows].abs().max().item() == 0.0
    assert _rel(out.float(), out2.float()) < 1.5e-2
    fin = torch.isfinite(lse2)
    assert torch.equal(torch.isfinite(lse), fin)
    assert (lse[fin] - lse2[fin]).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,L,Hq,Hkv,causal,pad", [
    (2, 256, 2, 2, True, "none"), (2, 128, 2, 2, True, "none"), (3, 200, 4, 4, True, "right"), (2, 96, 4, 4, True, "left"),
    (1, 300, 4, 2, True, "right"), (2, 384, 2, 2, False, "right"),
])
def test_attention_tc_backward(cuda_dev, B, L, Hq, Hkv, causal, pad):
    from dalm_b200 import ops
    D = 128
    torch.manual_seed(B * 77 + L)
    dev = cuda_dev
    qkv = torch.randn(B * L, (Hq + 2 * Hkv) * D, device=dev).to(bf16)
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    if pad == "right":
        for b in range(B): mask[b, L - 3 - 5 * b:] = 0
    elif pad == "left":
        for b in range(B): mask[b, :4 + 3 * b] = 0
    out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, Hq, Hkv, D, causal)
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qd, kd, vd, mask, causal, B, L, Hq, Hkv, D)
    rows = mask.bool().view(-1) if (causal and pad == "left") else torch.ones(B * L, dtype=torch.bool, device=dev)
    d_out = torch.randn(B * L, Hq * D, device=dev).to(bf16)
    d_out[~rows] = 0
    ref.backward(d_out.double())
    dqkv = torch.zeros(B * L, (Hq + 2 * Hkv) * D + 64, device=dev, dtype=bf16)          # outputs are column slices of a wider buffer
    dq, dk, dv = ops.attention_tc_bwd(q, k, v, mask, out, lse, d_out, B, L, Hq, Hkv, D, causal, dq=dqkv[:, :Hq * D],
                                      dk=dqkv[:, Hq * D:(Hq + Hkv) * D], dv=dqkv[:, (Hq + Hkv) * D:(Hq + 2 * Hkv) * D])
    assert _rel(dq.float(), qd.grad) < 3e-2
    assert _rel(dk.float(), kd.grad) < 3e-2
    assert _rel(dv.float(), vd.grad) < 3e-2
    assert dqkv[:, (Hq + 2 * Hkv) * D:].abs().max().item() == 0
    # and against the mma.sync kernels
    dq2, dk2, dv2 = ops.attention_bwd(q, k, v, mask, out, lse, d_out, B, L, Hq, Hkv, D, causal)
    assert _rel(dq.float(), dq2.float()) < 2e-2 and _rel(dk.float(), dk2.float()) < 2e-2 and _rel(dv.float(), dv2.float()) < 2e-2


def test_attention_tc_speed_report(cuda_dev, capsys):
    """not an assertion on speed, just a printed comparison at the cfg-3 decoder shape"""
    from dalm_b200 import ops
    B, L, H, D = 18, 256, 32, 128
    dev = cuda_dev
    qkv = torch.randn(B * L, 3 * H * D, device=dev).to(bf16)
    q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)

    def t(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n * 1e3
    out, lse = ops.attention_tc_fwd(q, k, v, mask, B, L, H, H, D, True)
    do = torch.randn_like(out); dq = torch.empty_like(out); dk = torch.empty_like(out); dv = torch.empty_like(out)
    res = {"fwd_tc_us": t(lambda: ops.attention_tc_fwd(q, k, v, mask, B, L, H, H, D, True, out=out)),
           "fwd_mma_us": t(lambda: ops.attention_fwd(q, k, v, mask, B, L, H, H, D, True, out=out)),
           "bwd_tc_us": t(lambda: ops.attention_tc_bwd(q, k, v, mask, out, lse, do, B, L, H, H, D, True, dq=dq, dk=dk, dv=dv)),
           "bwd_mma_us": t(lambda: ops.attention_bwd(q, k, v, mask, out, lse, do, B, L, H, H, D, True, dq=dq, dk=dk, dv=dv))}
    with capsys.disabled():
        print("\nATTN_TC_TIMING", res)
