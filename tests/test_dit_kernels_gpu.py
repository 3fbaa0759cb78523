"""Parity of the DiT kernels (tcgen05 GEMM / attention and the fused norm kernels) with the oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_dit

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


def g(seed):
    return torch.Generator().manual_seed(seed)


def bf16_close(got, want, ulps=1):
    """bf16 results of fp32-accumulated sums: identical up to accumulation order, i.e. all but a
    small fraction of elements are bit-equal and the rest are one bf16 ulp apart."""
    got, want = got.float().cpu(), want.float().cpu()
    diff = (got - want).abs()
    ulp = want.abs().clamp_min(1e-3) * 2.0 ** -7
    assert (diff <= ulp * (ulps + 0.01) + 1e-6).all(), f"max diff {diff.max().item()} (> 1 bf16 ulp) at {diff.argmax().item()}"
    frac = (diff > 0).float().mean().item()
    assert frac < 0.2, f"{frac:.3f} of the elements differ"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 192), (257, 1280, 1280), (1000, 768, 144), (64, 32, 64)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_bf16(cuda, M, N, K, epi):
    from worldforge_b200 import lib
    a = (torch.randn(M, K, generator=g(1)) * 0.5).to(BF)
    w = (torch.randn(N, K, generator=g(2)) * 0.1).to(BF)
    b = (torch.randn(N, generator=g(3)) * 0.1).to(BF)
    y = F.linear(a.float(), w.float(), b.float()).to(BF)        # one rounding of the fp32 sum
    if epi == lib.EPI_BF16:
        out = torch.empty(M, N, dtype=BF, device=cuda)
        lib.gemm_bf16(a.to(cuda), w.to(cuda), b.to(cuda), out, epi)
        bf16_close(out, y)
    elif epi == lib.EPI_GELU_BF16:
        out = torch.empty(M, N, dtype=BF, device=cuda)
        lib.gemm_bf16(a.to(cuda), w.to(cuda), b.to(cuda), out, epi)
        want = F.gelu(y, approximate="tanh")
        torch.testing.assert_close(out.cpu().float(), want.float(), rtol=2e-2, atol=2e-3)
    elif epi == lib.EPI_RESID_F32:
        x = torch.randn(M, N, generator=g(4))
        gate = torch.randn(N, generator=g(5))
        out = x.clone().to(cuda)
        lib.gemm_bf16(a.to(cuda), w.to(cuda), b.to(cuda), out, epi, gate=gate.to(cuda))
        want = x + y * gate
        torch.testing.assert_close(out.cpu(), want, rtol=1e-2, atol=2e-2)
        out2 = x.clone().to(cuda)
        lib.gemm_bf16(a.to(cuda), w.to(cuda), None, out2, epi)       # no bias, no gate: x + y
        want2 = x + F.linear(a.float(), w.float()).to(BF)
        torch.testing.assert_close(out2.cpu(), want2, rtol=1e-2, atol=2e-2)
    else:
        out = torch.empty(M, N, dtype=F32, device=cuda)
        lib.gemm_bf16(a.to(cuda), w.to(cuda), b.to(cuda), out, epi)
        bf16_close(out, y)


def test_gemm_strided_views(cuda):
    """A and the output may be column slices of wider matrices (fused QKV buffer)."""
    from worldforge_b200 import lib
    M, N, K = 200, 256, 128
    big_a = (torch.randn(M, 3 * K, generator=g(6)) * 0.5).to(BF).to(cuda)
    w = (torch.randn(N, K, generator=g(7)) * 0.1).to(BF).to(cuda)
    big_o = torch.zeros(M, 2 * N, dtype=BF, device=cuda)
    lib.gemm_bf16(big_a[:, K:2 * K], w, None, big_o[:, N:], lib.EPI_BF16)
    want = F.linear(big_a[:, K:2 * K].float().cpu(), w.float().cpu()).to(BF)
    bf16_close(big_o[:, N:], want)
    assert (big_o[:, :N] == 0).all()


@pytest.mark.parametrize("Lq,Lk,heads", [(256, 64, 1), (300, 300, 2), (700, 257, 2), (130, 512, 1), (1000, 1000, 3)])
def test_attention(cuda, Lq, Lk, heads):
    from worldforge_b200 import lib
    D = heads * 128
    q = torch.randn(Lq, D, generator=g(8)).to(BF)
    k = torch.randn(Lk, D, generator=g(9)).to(BF)
    v = torch.randn(Lk, D, generator=g(10)).to(BF)
    want = wan_dit.attention(q.view(Lq, heads, 128), k.view(Lk, heads, 128), v.view(Lk, heads, 128), amp=True).reshape(Lq, D)
    out = torch.empty(Lq, D, dtype=BF, device=cuda)
    lib.attention_bf16(q.to(cuda), k.to(cuda), v.to(cuda), out, heads)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=2e-2, atol=4e-3)


def test_attention_peaked_rows_trigger_rescale(cuda):
    """Scores that grow along the key axis force the lazy O-rescale path."""
    from worldforge_b200 import lib
    Lq, Lk = 256, 640
    q = torch.randn(Lq, 128, generator=g(11))
    k = torch.randn(Lk, 128, generator=g(12)) * (1.0 + 3.0 * torch.arange(Lk).float().unsqueeze(1) / Lk)
    v = torch.randn(Lk, 128, generator=g(13))
    q, k, v = q.to(BF), k.to(BF), v.to(BF)
    want = wan_dit.attention(q.view(Lq, 1, 128), k.view(Lk, 1, 128), v.view(Lk, 1, 128), amp=True).reshape(Lq, 128)
    out = torch.empty(Lq, 128, dtype=BF, device=cuda)
    lib.attention_bf16(q.to(cuda), k.to(cuda), v.to(cuda), out, 1)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=3e-2, atol=6e-3)


def test_attention_add_in_and_strided(cuda):
    """Cross-attention form: q/k/v are slices of fused buffers and the image branch is added (model.py:227)."""
    from worldforge_b200 import lib
    Lq, Lk, heads = 384, 257, 2
    D = heads * 128
    qkv = torch.randn(Lq, 3 * D, generator=g(14)).to(BF).to(cuda)
    kv = torch.randn(Lk, 2 * D, generator=g(15)).to(BF).to(cuda)
    prev = torch.randn(Lq, D, generator=g(16)).to(BF).to(cuda)
    out = torch.empty(Lq, D, dtype=BF, device=cuda)
    lib.attention_bf16(qkv[:, :D], kv[:, :D], kv[:, D:], out, heads, add_in=prev)
    a = wan_dit.attention(qkv[:, :D].cpu().view(Lq, heads, 128), kv[:, :D].cpu().view(Lk, heads, 128),
                          kv[:, D:].cpu().view(Lk, heads, 128), amp=True).reshape(Lq, D)
    want = (a.float() + prev.cpu().float()).to(BF)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=2e-2, atol=1e-2)


@pytest.mark.parametrize("rows,D", [(37, 256), (300, 5120), (5, 1280), (700, 5120), (1500, 256)])   # >= 592 rows: the persistent row-ring form
def test_layer_norm_modulate(cuda, rows, D):
    from worldforge_b200 import lib
    x = torch.randn(rows, D, generator=g(17)) * 2 + 0.3
    sc, sh = torch.randn(D, generator=g(18)) * 0.1, torch.randn(D, generator=g(19)) * 0.1
    want = (wan_dit.layer_norm(x, 1e-6).float() * (1 + sc) + sh).to(BF)
    out = torch.empty(rows, D, dtype=BF, device=cuda)
    lib.layer_norm(x.to(cuda), out, 1e-6, scale=sc.to(cuda), shift=sh.to(cuda))
    bf16_close(out, want)
    # block 0: the token stream is bf16 and LN's result is rounded to bf16 before modulation
    xb = x.to(BF)
    want0 = (wan_dit.layer_norm(xb, 1e-6).float() * (1 + sc) + sh).to(BF)
    out0 = torch.empty(rows, D, dtype=BF, device=cuda)
    lib.layer_norm(xb.float().to(cuda), out0, 1e-6, scale=sc.to(cuda), shift=sh.to(cuda), round_norm_bf16=True)
    # here LN's value is rounded to bf16 BEFORE the modulation: where that rounding is a near-tie the two
    # implementations may pick neighbouring bf16 values, which (1+scale) then carries into the result
    d0 = (out0.float().cpu() - want0.float()).abs()
    assert (d0 > 0).float().mean() < 5e-3 and d0.max() <= 2.0 ** -5, (float((d0 > 0).float().mean()), float(d0.max()))
    # affine (norm3) and bf16 input / fp32 output variants
    w, b = 1 + torch.randn(D, generator=g(20)) * 0.1, torch.randn(D, generator=g(21)) * 0.1
    want3 = F.layer_norm(x, (D,), w, b, 1e-6)
    out3 = torch.empty(rows, D, dtype=F32, device=cuda)
    lib.layer_norm(x.to(cuda), out3, 1e-6, weight=w.to(cuda), bias=b.to(cuda))
    torch.testing.assert_close(out3.cpu(), want3, rtol=1e-5, atol=1e-5)
    outb = torch.empty(rows, D, dtype=BF, device=cuda)
    lib.layer_norm(xb.to(cuda), outb, 1e-5, weight=w.to(cuda), bias=b.to(cuda))
    bf16_close(outb, F.layer_norm(xb.float(), (D,), w, b, 1e-5).to(BF))


@pytest.mark.parametrize("grid,heads", [((2, 3, 4), 2), ((3, 30, 52), 1)])
def test_rms_norm_rope(cuda, grid, heads):
    from worldforge_b200 import lib, transformer
    L, D = grid[0] * grid[1] * grid[2], heads * 128
    big = torch.randn(L, 3 * D, generator=g(22)).to(BF)
    w = 1 + torch.randn(D, generator=g(23)) * 0.1
    q = big[:, D:2 * D]
    n = wan_dit.rms_norm(q, w, 1e-6)
    want_rope = wan_dit.rope_apply(n.view(L, heads, 128), grid).reshape(L, D).to(BF)
    want_plain = n.to(BF)
    rope = transformer.rope_table(grid).to(cuda)
    buf = big.clone().to(cuda)
    lib.rms_norm_rope_(buf[:, D:2 * D], w.to(cuda), 1e-6, rope)
    bf16_close(buf[:, D:2 * D], want_rope)
    assert torch.equal(buf[:, :D].cpu(), big[:, :D]) and torch.equal(buf[:, 2 * D:].cpu(), big[:, 2 * D:])
    buf2 = big.clone().to(cuda)
    lib.rms_norm_rope_(buf2[:, D:2 * D], w.to(cuda), 1e-6, None)
    bf16_close(buf2[:, D:2 * D], want_plain)


def test_patchify_head_gemv(cuda):
    from worldforge_b200 import lib
    cfg = wan_dit.DitConfig(dim=256, num_heads=2)
    x = torch.randn(36, 3, 8, 12, generator=g(24)).to(BF)
    want, grid = wan_dit.patchify(x, cfg)
    cols = torch.empty(want.shape, dtype=BF, device=cuda)
    lib.patchify(x.to(cuda), cols)
    assert torch.equal(cols.cpu(), want)
    L, D = want.shape[0], 256
    tok = torch.randn(L, D, generator=g(25))
    sc, sh = torch.randn(D, generator=g(26)) * 0.1, torch.randn(D, generator=g(27)) * 0.1
    hw, hb = torch.randn(64, D, generator=g(28)) * 0.05, torch.randn(64, generator=g(29)) * 0.05
    h = wan_dit.layer_norm(tok, 1e-6) * (1 + sc) + sh
    want_o = wan_dit.unpatchify(F.linear(h, hw, hb), cfg, grid)
    out = torch.empty(16, 3, 8, 12, dtype=F32, device=cuda)
    lib.dit_head(tok.to(cuda), sc.to(cuda), sh.to(cuda), hw.to(cuda), hb.to(cuda), out, grid, 1e-6)
    torch.testing.assert_close(out.cpu(), want_o, rtol=1e-4, atol=1e-5)
    W = torch.randn(512, 256, generator=g(30)) * 0.05
    xv, bv = torch.randn(256, generator=g(31)), torch.randn(512, generator=g(32))
    o = torch.empty(512, device=cuda)
    lib.gemv_f32(W.to(cuda), xv.to(cuda), bv.to(cuda), o, silu_in=True, silu_out=False)
    torch.testing.assert_close(o.cpu(), F.linear(F.silu(xv), W, bv), rtol=1e-4, atol=1e-5)
    lib.gemv_f32(W.to(cuda), xv.to(cuda), bv.to(cuda), o, silu_in=False, silu_out=True)
    torch.testing.assert_close(o.cpu(), F.silu(F.linear(xv, W, bv)), rtol=1e-4, atol=1e-5)
    t = torch.randn(1000, generator=g(33)).to(BF)
    tg = t.clone().to(cuda)
    lib.gelu_erf_bf16_(tg)
    torch.testing.assert_close(tg.cpu().float(), F.gelu(t).float(), rtol=1e-2, atol=1e-3)
