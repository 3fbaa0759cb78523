"""Parity at the shapes of BASELINE.json's configurations (VERDICT r1 "missing #1"): Wan2.1-I2V-14B widths
(dim 5120, 40 heads of 128, ffn 13 824), the 480p / 81-frame token count L = 32 760 (C2), the 720p count 75 600 (C3),
the 9-frame count 4 680 (C1), the N=8 token shard M = 4 095, and full-resolution VAE layers.

Tolerances are MEASURED FLOORS, not constants: for every floating-point path three results are formed from the same
inputs -
    truth   the oracle in fp32 (no bf16 / tf32 rounding anywhere),
    model   the oracle with the reference's GPU rounding points (bf16 autocast for the DiT, wan_dit ``amp=True``;
            tf32 convolution operands for the VAE, cuDNN's default for the reference's fp32 VAE),
    engine  the CUDA path through the C ABI -
and the assertion is  err(engine, truth) <= 1.1 * err(model, truth):  the engine may not be further from the exact
answer than the reference's own arithmetic is.  The measured floors are printed (pytest -s shows them; the GPU session
logs are committed under profiles/).  Where the oracle is too slow for the full size on the CPU it is evaluated on the
GPU (the oracle functions are plain torch and device-agnostic; TF32 matmul/conv is switched OFF for it) and tied to
its CPU evaluation on a subset of rows.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_dit, wan_vae

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32
FLOOR = 1.1          # engine error <= FLOOR x the error of the reference's own rounding points


def g(seed, device="cpu"):
    return torch.Generator(device=device).manual_seed(seed)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(autouse=True)
def exact_fp32_reference():
    """The fp32 references below must be fp32: no TF32 in torch's own matmuls / convolutions."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def bf16_close_gpu(got, want, ulps=1, max_frac=0.2, atol=1e-6):
    """bf16 results of fp32-accumulated sums agree up to accumulation order: bit-equal except for a small fraction of
    elements one bf16 ulp apart (same criterion as tests/test_dit_kernels_gpu.py, evaluated on the device).  ``atol`` is
    the fp32 accumulation error itself - for sums that cancel to near zero it, not the bf16 spacing, is what two
    summation orders can differ by."""
    got, want = got.float(), want.float()
    diff = (got - want).abs()
    ulp = want.abs().clamp_min(1e-3) * 2.0 ** -7
    bad = (diff > ulp * (ulps + 0.01) + atol)
    assert not bool(bad.any()), (f"{int(bad.sum())} elements beyond {ulps} bf16 ulp + {atol:.1e}; worst "
                                 f"{(diff - ulp * (ulps + 0.01)).max().item():.3e} over the bound")
    frac = (diff > 0).float().mean().item()
    assert frac < max_frac, f"{frac:.3f} of the elements differ"
    return frac


def accumulation_atol(a, w, K):
    """Bound on the difference of two fp32 summation orders of sum_k a_k w_k: a few units of 2^-24 of sum_k |a_k w_k|."""
    return 16.0 * 2.0 ** -24 * K * a.float().abs().mean().item() * w.float().abs().mean().item()


# ----------------------------------------------------------------------------------------------------------------
# (i) self-attention at the full token counts
# ----------------------------------------------------------------------------------------------------------------

def oracle_attention_on_device(q, k, v, heads, amp, chunk=2048):
    """wan_dit.attention (reference attention.py:24-130 restated) over all query rows, evaluated on the tensors' device
    in query chunks (each row's softmax still sees every key at once)."""
    L, Lk = q.shape[0], k.shape[0]
    kk, vv = k.view(Lk, heads, 128), v.view(Lk, heads, 128)
    out = torch.empty(L, heads * 128, dtype=F32, device=q.device)
    for s in range(0, L, chunk):
        e = min(L, s + chunk)
        out[s:e] = wan_dit.attention(q[s:e].view(e - s, heads, 128), kk, vv, amp=amp).reshape(e - s, -1).float()
    return out


@pytest.mark.parametrize("L,heads", [(32760, 2), (75600, 1), (4680, 3)])
def test_self_attention_full_sequence(cuda, L, heads):
    from worldforge_b200 import lib
    D = heads * 128
    gd = g(100 + heads, "cuda")
    q = torch.randn(L, D, generator=gd, device=cuda).to(BF)
    k = torch.randn(L, D, generator=gd, device=cuda).to(BF)
    v = torch.randn(L, D, generator=gd, device=cuda).to(BF)
    out = torch.empty(L, D, dtype=BF, device=cuda)
    lib.attention_bf16(q, k, v, out, heads)
    truth = oracle_attention_on_device(q, k, v, heads, amp=False)
    model = oracle_attention_on_device(q, k, v, heads, amp=True)
    e_model, e_engine, e_em = rel(model, truth), rel(out, truth), rel(out, model)
    print(f"\n[floor] self-attention L={L} heads={heads}: model-vs-fp32 {e_model:.3e}  engine-vs-fp32 {e_engine:.3e}  "
          f"engine-vs-model {e_em:.3e}")
    assert e_engine <= FLOOR * e_model, (e_engine, e_model)
    # engine and model round the probabilities to bf16 at different scales (the kernel's lazily updated running maximum
    # and exp2 against the oracle's row maximum and exp), so their rounding errors are independent: the distance between
    # them is at most sqrt(2) x the floor, not below it
    assert e_em <= 2 ** 0.5 * FLOOR * e_model, (e_em, e_model)
    # every row, including the last (partial) query tile and the masked tail of the last key block
    row_err = (out.float() - truth).norm(dim=1) / truth.norm(dim=1)
    row_floor = (model - truth).norm(dim=1) / truth.norm(dim=1)
    assert bool((row_err <= 2.0 * row_floor.max()).all()), (row_err.max().item(), row_floor.max().item())
    # the device-evaluated oracle is the CPU oracle: first rows, a middle stretch, the ragged tail
    rows = torch.cat([torch.arange(0, 192), torch.arange(L // 2 - 64, L // 2 + 64), torch.arange(L - 200, L)])
    cpu = wan_dit.attention(q[rows.to(cuda)].cpu().view(-1, heads, 128), k.cpu().view(L, heads, 128),
                            v.cpu().view(L, heads, 128), amp=True).reshape(len(rows), D).float()
    tie = rel(model[rows.to(cuda)].cpu(), cpu)
    assert tie <= 0.25 * e_model, (tie, e_model)     # same function up to summation order
    assert rel(out[rows.to(cuda)].cpu(), cpu) <= 2 ** 0.5 * FLOOR * e_model


# ----------------------------------------------------------------------------------------------------------------
# (ii) one full-width Wan2.1-14B forward (patch embedding, time / text / image embedders, ONE block, head) at C1's tokens
# ----------------------------------------------------------------------------------------------------------------

def test_full_width_forward_one_block(cuda):
    from worldforge_b200.transformer import WanDitConfig, WfWanTransformer
    cfg = wan_dit.DitConfig(num_layers=1)                        # dim 5120, 40 heads, ffn 13824 (wan_i2v_14B.py:27-36)
    P = wan_dit.init_params(cfg, 5)
    grid = (3, 60, 104)                                          # 480p, 9 frames -> L = 3*30*52 = 4680 (config 1)
    gg = g(0)
    x = torch.randn(1, 36, *grid, generator=gg).to(BF)
    ctx = torch.randn(1, 512, 4096, generator=gg).to(BF)
    ctx[:, 64:] = 0                                              # padded prompt tail (pipeline_wan_i2v_clean.py:196-199)
    clip = torch.randn(1, 257, 1280, generator=gg).to(BF)
    t = torch.tensor([737])
    with torch.no_grad():
        model = wan_dit.dit_forward(P, cfg, x[0], t, ctx[0], clip[0], amp=True).to(BF).float()
        truth = wan_dit.dit_forward(P, cfg, x[0], t, ctx[0], clip[0], amp=False)
    m = WfWanTransformer.from_state_dict(P, WanDitConfig(num_layers=1), cuda)
    got = m(x.to(cuda), t.to(cuda), ctx.to(cuda), clip.to(cuda), return_dict=False)[0][0].float().cpu()
    e_model, e_engine, e_em = rel(model, truth), rel(got, truth), rel(got, model)
    print(f"\n[floor] Wan-14B-width forward, 1 block, L=4680: model-vs-fp32 {e_model:.3e}  engine-vs-fp32 {e_engine:.3e}  "
          f"engine-vs-model {e_em:.3e}")
    assert e_engine <= FLOOR * e_model, (e_engine, e_model)
    assert e_em <= e_model


# ----------------------------------------------------------------------------------------------------------------
# (iii) the block's GEMMs at K = 5120 / 13824, M = 4095 (N=8 token shard) and M = 32760
# ----------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("M,N,K,epi", [
    (4095, 5120, 5120, 0), (4095, 15360, 5120, 0), (4095, 13824, 5120, 1), (4095, 5120, 13824, 2),
    (32760, 15360, 5120, 0), (32760, 5120, 13824, 2), (32760, 13824, 5120, 1), (4680, 5120, 5120, 3)])
def test_gemm_real_shapes(cuda, M, N, K, epi):
    from worldforge_b200 import lib
    gd = g(7, "cuda")
    a = (torch.randn(M, K, generator=gd, device=cuda) * 0.5).to(BF)
    w = (torch.randn(N, K, generator=gd, device=cuda) * 0.02).to(BF)
    b = (torch.randn(N, generator=gd, device=cuda) * 0.02).to(BF)
    y32 = torch.empty(M, N, dtype=F32, device=cuda)
    for s in range(0, M, 8192):                                  # fp32 SGEMM (TF32 off) of the bf16 operands, one rounding
        y32[s:s + 8192] = F.linear(a[s:s + 8192].float(), w.float(), b.float())
    y = y32.to(BF)
    atol = accumulation_atol(a, w, K)
    if epi == lib.EPI_BF16:
        out = torch.empty(M, N, dtype=BF, device=cuda)
        lib.gemm_bf16(a, w, b, out, epi)
        frac = bf16_close_gpu(out, y, atol=atol)
    elif epi == lib.EPI_GELU_BF16:
        out = torch.empty(M, N, dtype=BF, device=cuda)
        lib.gemm_bf16(a, w, b, out, epi)
        want = F.gelu(y, approximate="tanh").float()
        # y itself may differ by one bf16 ulp (or the accumulation error) between two summation orders; GELU carries that
        # to its output with slope |gelu'(y)|, which in the negative tail is many OUTPUT ulps: bound = slope x input
        # freedom + 2 output ulps (tanhf vs torch's tanh, output rounding)
        yf = y.float().requires_grad_(True)
        slope = torch.autograd.grad(F.gelu(yf, approximate="tanh").sum(), yf)[0].abs()
        tol = slope * (y.float().abs() * 2.0 ** -7 * 1.01 + atol) + want.abs().clamp_min(1e-3) * 2.0 ** -7 * 2.01 + 1e-6
        d = (out.float() - want).abs()
        assert bool((d <= tol).all()), f"{int((d > tol).sum())} elements beyond the bound; worst {(d - tol).max().item():.3e}"
        frac = (d > 0).float().mean().item()
        assert frac < 0.2, frac
    elif epi == lib.EPI_RESID_F32:
        x = torch.randn(M, N, generator=gd, device=cuda)
        gate = torch.randn(N, generator=gd, device=cuda)
        out = x.clone()
        lib.gemm_bf16(a, w, b, out, epi, gate=gate)
        want = x + y.float() * gate
        # identical fp32 expression; the only freedom is the bf16 rounding of y where the fp32 sums differ in their last bits
        d = (out - want).abs()
        tol = (y.float().abs().clamp_min(1e-3) * 2.0 ** -7 * 1.01 + atol) * gate.abs() + 1e-6
        assert bool((d <= tol).all()), f"{int((d > tol).sum())} elements beyond the bound; worst {(d - tol).max().item():.3e}"
        frac = (d > 1e-7).float().mean().item()
        assert frac < 0.2
    else:
        out = torch.empty(M, N, dtype=F32, device=cuda)
        lib.gemm_bf16(a, w, b, out, epi)
        frac = bf16_close_gpu(out, y, atol=atol)
    print(f"\n[gemm] M={M} N={N} K={K} epi={epi}: {frac:.4f} of the outputs differ from the fp32-accumulated reference (<= 1 bf16 ulp)")
    if M == 4095 and N == 5120 and K == 5120:                    # and the CPU oracle's statement of the same Linear
        rows = torch.cat([torch.arange(0, 64), torch.arange(4095 - 31, 4095)])
        want = wan_dit.amp_linear(a[rows.to(cuda)].cpu(), w.cpu().float(), b.cpu().float(), amp=True)
        bf16_close_gpu(out[rows.to(cuda)].cpu(), want, atol=atol)


# ----------------------------------------------------------------------------------------------------------------
# (iv) VAE convolutions at full resolution
# ----------------------------------------------------------------------------------------------------------------

def tf32_trunc(t):
    return (t.contiguous().view(torch.int32) & -8192).view(F32)


def tf32_rna(t):
    return ((t.contiguous().view(torch.int32) + 4096) & -8192).view(F32)


@pytest.mark.parametrize("C,T,H,W", [(96, 21, 480, 832), (384, 21, 60, 104), (192, 11, 240, 416)])
def test_vae_conv333_full_resolution(cuda, C, T, H, W):
    """One 3x3x3 causal convolution of the decoder at its real extent: the 96-channel layers at 480x832 (21 of the
    81 frames - the whole clip at once is 3.1 G elements, beyond cuDNN's 2^31 indexing for the reference), the
    384-channel layers at 60x104 and the 192-channel ones at 240x416."""
    from worldforge_b200 import lib, vae as wvae
    gd = g(9, "cuda")
    x = torch.randn(T, H, W, C, generator=gd, device=cuda)                      # channels-last, as the engine holds it
    w = torch.randn(C, C, 3, 3, 3, generator=gd, device=cuda) / (C * 27) ** 0.5
    b = torch.randn(C, generator=gd, device=cuda) * 0.1
    tile_w = 16 if W % 16 == 0 else 8
    out = torch.empty(T, H, W, C, device=cuda)
    lib.conv_tf32(x, wvae._w_conv3d(w), b, wvae.TAPS_333, out, T=T, H=H, W=W, Cout=C, tile_w=tile_w)
    xp = x.permute(3, 0, 1, 2)                                                  # [C, T, H, W] view for the oracle
    ref = lambda xx, ww: wan_vae.causal_conv3d(xx, ww, b).permute(1, 2, 3, 0)   # oracle/wan_vae.py on the device, fp32
    truth = ref(xp, w)
    e_raw = rel(out, truth)
    e_trunc = rel(ref(tf32_trunc(xp), tf32_trunc(w)), truth)                    # tf32 operands by truncation (the tensor core's own)
    e_rna = rel(ref(tf32_rna(xp), tf32_rna(w)), truth)                          # tf32 operands rounded to nearest
    torch.backends.cudnn.allow_tf32 = True                                      # the reference's default on a GPU
    e_cudnn = rel(ref(xp, w), truth)
    torch.backends.cudnn.allow_tf32 = False
    # the way the engine calls it: operands rounded to tf32 by their producers (weights at load, activations by the
    # RMS-norm / layout kernels), so that the tensor core's truncation is exact and the arithmetic is cuDNN's
    lib.conv_tf32(lib.round_tf32(x), wvae._round_tf32_host(wvae._w_conv3d(w)), b, wvae.TAPS_333, out, T=T, H=H, W=W, Cout=C,
                  tile_w=tile_w)
    e_engine = rel(out, truth)
    print(f"\n[floor] conv3x3x3 C={C} {T}x{H}x{W}: engine-vs-fp32 {e_engine:.3e} (raw fp32 operands: {e_raw:.3e}); "
          f"tf32-truncated operands {e_trunc:.3e}, tf32-rounded {e_rna:.3e}, cuDNN allow_tf32 (reference default) {e_cudnn:.3e}")
    assert e_raw <= FLOOR * e_trunc, (e_raw, e_trunc)
    assert e_engine <= FLOOR * e_cudnn, (e_engine, e_cudnn)
    # stored rounded (round_out): equal to rounding the plain result
    out2 = torch.empty_like(out)
    lib.conv_tf32(lib.round_tf32(x), wvae._round_tf32_host(wvae._w_conv3d(w)), b, wvae.TAPS_333, out2, T=T, H=H, W=W, Cout=C,
                  tile_w=tile_w, round_out=True)
    assert torch.equal(out2, tf32_rna(out))


def test_vae_resblock_full_resolution(cuda):
    """ResidualBlock (vae.py:186-220) of the last decoder stage, 96 channels at 480x832, 5 frames: RMS-norm + SiLU,
    two 3x3x3 convolutions, residual - engine vs oracle/wan_vae.res_block evaluated in fp32 and with cuDNN's tf32."""
    from worldforge_b200 import vae as wvae
    cfg = wan_vae.VaeConfig()
    P = wan_vae.init_params(cfg, 11, device=cuda)
    m = wvae.WfWanVAE(P, cuda)
    name = next(n for kind, n, cin, cout in m.dec_plan[::-1] if kind == "res" and cin == cout == 96)
    x = torch.randn(96, 5, 480, 832, generator=g(12, "cuda"), device=cuda)
    truth = wan_vae.res_block(P, name, x)
    torch.backends.cudnn.allow_tf32 = True
    model = wan_vae.res_block(P, name, x)
    torch.backends.cudnn.allow_tf32 = False
    got = m._res(x.permute(1, 2, 3, 0).contiguous(), name, 96, 96)[0].permute(3, 0, 1, 2)
    e_model, e_engine = rel(model, truth), rel(got, truth)
    print(f"\n[floor] VAE ResidualBlock 96ch 5x480x832: cuDNN-tf32-vs-fp32 {e_model:.3e}  engine-vs-fp32 {e_engine:.3e}")
    assert e_engine <= FLOOR * e_model, (e_engine, e_model)


def test_vae_full_width_round_trip_9_frames(cuda):
    """The whole 96-wide VAE (encode and decode) on a 9-frame 480x832 clip - config 1's extent - against the oracle
    evaluated on the device in fp32 (truth) and with cuDNN's tf32 convolutions (the reference's default)."""
    from worldforge_b200 import vae as wvae
    cfg = wan_vae.VaeConfig()
    P = wan_vae.init_params(cfg, 13, device=cuda)
    m = wvae.WfWanVAE(P, cuda)
    video = torch.rand(3, 9, 480, 832, generator=g(14, "cuda"), device=cuda) * 2 - 1
    z = torch.randn(16, 3, 60, 104, generator=g(15, "cuda"), device=cuda)
    with torch.no_grad():
        t_mu, t_dec = wan_vae.encode_mode(P, cfg, video), wan_vae.decode(P, cfg, z)
        torch.backends.cudnn.allow_tf32 = True
        m_mu, m_dec = wan_vae.encode_mode(P, cfg, video), wan_vae.decode(P, cfg, z)
        torch.backends.cudnn.allow_tf32 = False
    g_mu = m.encode(video.unsqueeze(0)).latent_dist.mode()[0]
    g_dec = m.decode(z.unsqueeze(0))[0][0]
    for what, got, model, truth in (("encode", g_mu, m_mu, t_mu), ("decode", g_dec, m_dec, t_dec)):
        e_model, e_engine = rel(model, truth), rel(got, truth)
        print(f"\n[floor] VAE {what} 9x480x832, dim 96: cuDNN-tf32-vs-fp32 {e_model:.3e}  engine-vs-fp32 {e_engine:.3e}")
        assert e_engine <= 1.25 * e_model, (what, e_engine, e_model)   # + the pre-summed upsample / space-to-depth weight forms
