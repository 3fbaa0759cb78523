"""LongCat 720p refine pass on the engine vs the oracle: input upsampling, the DiT with block-sparse
self-attention, LoRA folding, and the refine loop (longcat_video/pipeline_longcat_video.py:1271-1511)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import adapters, longcat_dit as old, longcat_sched as ols, wan_vae

pytestmark = pytest.mark.gpu
BF = torch.bfloat16

DIT = dict(hidden_size=256, depth=2, num_heads=2, caption_channels=64, adaln_tembed_dim=32, frequency_embedding_size=32)
BSA = dict(sparsity=0.5, cdf_threshold=None, chunk_3d_shape_q=[4, 4, 4], chunk_3d_shape_k=[4, 4, 4])


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("double_frames", [False, True])
def test_refine_upsample_matches_torch_chain(cuda, double_frames):
    """wf_refine_upsample vs the reference's tensor expressions (:1396-1421) evaluated by torch on the same GPU."""
    from worldforge_b200 import lib
    Fr, H0, W0, H, W = 5, 24, 40, 36, 64
    vid = torch.randint(0, 256, (Fr, H0, W0, 3), generator=g(0), dtype=torch.uint8)
    F2 = 2 * Fr if double_frames else Fr
    s1 = vid.to(cuda).permute(0, 3, 1, 2).to(BF)
    down = F.interpolate(s1, size=(H, W), mode="bilinear", align_corners=True).permute(1, 0, 2, 3).unsqueeze(0) / 255.0
    up = F.interpolate(down, size=(F2, H, W), mode="trilinear", align_corners=True) * 2 - 1
    up = torch.cat([up[:, :, 0:1].repeat(1, 1, 3, 1, 1), up, up[:, :, -1:].repeat(1, 1, 2, 1, 1)], dim=2)
    got = lib.refine_upsample(vid.to(cuda), F2, H, W, 3, 2)
    assert got.shape == (3, F2 + 5, H, W) and got.dtype == torch.float32
    d = (got - up[0].float()).abs()
    # fp32 interpolation with a different FMA contraction, then bf16 rounding: rare one-ulp flips only (ulp <= 2^-8 on [-1,1])
    assert d.max().item() <= 2 ** -7 and (d > 0).float().mean().item() < 2e-2, (d.max().item(), (d > 0).float().mean().item())


def _model(cuda):
    from worldforge_b200 import longcat
    ocfg, pcfg = old.LongCatConfig(**DIT), longcat.LongCatConfig(**DIT)
    P = old.init_params(ocfg, 3)
    m = longcat.WfLongCatTransformer.from_state_dict(P, pcfg, cuda)
    m.bsa_params = dict(BSA)
    return P, ocfg, m


def test_dit_with_bsa_matches_oracle(cuda):
    P, ocfg, m = _model(cuda)
    x = torch.randn(1, 16, 8, 16, 16, generator=g(9))
    ts = torch.tensor([[0.0] * 4 + [600.0] * 4])
    ctx = torch.randn(1, 1, 10, 64, generator=g(1)).to(BF)
    mask = torch.ones(1, 10, dtype=torch.int64); mask[:, 8:] = 0
    run = lambda: m(x.to(cuda).to(BF), ts.to(cuda).to(BF), ctx.to(cuda), encoder_attention_mask=mask.to(cuda), num_cond_latents=4)[0].cpu()
    dense = run()
    m.enable_bsa()
    sparse = run()
    m.disable_bsa()
    want = old.dit_forward(P, ocfg, x[0].to(BF), ts[0].to(BF), ctx[0, 0, :8], num_cond_latents=4, amp=True, bsa=dict(BSA))
    want_dense = old.dit_forward(P, ocfg, x[0].to(BF), ts[0].to(BF), ctx[0, 0, :8], num_cond_latents=4, amp=True)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(dense, want_dense) < 8e-3
    # a chunk whose gating score sits within bf16 noise of the top-k threshold may be chosen differently; with 8 key chunks
    # per head that moves the output by less than the sparse-vs-dense difference itself
    assert rel(sparse, want) < 1.2e-2, rel(sparse, want)
    assert rel(sparse, want) < 0.5 * rel(want_dense, want) or rel(sparse, want) < 8e-3


def test_merge_lora_equals_side_branch(cuda):
    """W + s*up@down reproduces org(x) + s*up(down(x)) (longcat_video_dit.py:234-249) up to the branch's bf16 rounding."""
    from worldforge_b200 import longcat
    sd = {"blocks.0.attn.qkv.weight": torch.randn(96, 32, generator=g(1)) * 0.1,
          "blocks.0.attn.proj.weight": torch.randn(32, 32, generator=g(2)) * 0.1}
    H = "___lorahyphen___"
    n_qkv, n_proj = f"lora{H}blocks{H}0{H}attn{H}qkv", f"lora{H}blocks{H}0{H}attn{H}proj"
    lora = {n_qkv + ".lora_down.weight": torch.randn(3 * 4, 32, generator=g(3)) * 0.1,
            n_proj + ".lora_down.weight": torch.randn(4, 32, generator=g(4)) * 0.1,
            n_proj + ".lora_up.weight": torch.randn(32, 4, generator=g(5)) * 0.1}
    for i in range(3):
        lora[f"{n_qkv}.lora_up.blocks.{i}.weight"] = torch.randn(32, 4, generator=g(6 + i)) * 0.1
    merged = longcat.merge_lora(sd, lora, multiplier=1.0, rank=4, alpha=2.0)
    x = torch.randn(7, 32, generator=g(20))
    lx = F.linear(x, lora[n_qkv + ".lora_down.weight"])
    side = torch.cat([F.linear(lx[:, 4 * i:4 * i + 4], lora[f"{n_qkv}.lora_up.blocks.{i}.weight"]) for i in range(3)], dim=-1)
    torch.testing.assert_close(F.linear(x, merged["blocks.0.attn.qkv.weight"]), F.linear(x, sd["blocks.0.attn.qkv.weight"]) + 0.5 * side,
                               rtol=1e-5, atol=1e-6)
    side = F.linear(F.linear(x, lora[n_proj + ".lora_down.weight"]), lora[n_proj + ".lora_up.weight"])
    torch.testing.assert_close(F.linear(x, merged["blocks.0.attn.proj.weight"]), F.linear(x, sd["blocks.0.attn.proj.weight"]) + 0.5 * side,
                               rtol=1e-5, atol=1e-6)


def test_refine_schedule_and_loop_match_oracle(cuda):
    """6 refine steps (t_thresh 0.6 of a 10-step schedule) with BSA on: engine loop + scheduler vs the oracle's."""
    from worldforge_b200 import longcat_pipeline as lp
    P, ocfg, m = _model(cuda)
    m.enable_bsa()
    lat0 = torch.randn(1, 16, 8, 16, 16, generator=g(21))
    pe = torch.randn(1, 1, 8, 64, generator=g(22)).to(BF)
    pm = torch.ones(1, 8, dtype=torch.int64); pm[0, 6:] = 0
    so = ols.OracleEuler(1000, 1.0)
    ts_o = ols.refine_schedule(so, 10, 0.6)
    want = ols.refine_loop(adapters.OracleLongCatDit(P, ocfg, amp=True, bsa=dict(BSA)), so, lat0.clone(), pe, pm, 4, ts_o)
    se = lp.WfFlowMatchEulerScheduler(1000, 1.0)
    ts_e = lp.refine_schedule(se, 10, 0.6, device=cuda)
    assert torch.equal(ts_e.cpu(), ts_o)
    # sigmas = timesteps / 1000 is evaluated where the timesteps live: on the GPU (as in the reference run) torch's kernel
    # multiplies by the fp32 reciprocal of a host scalar, the CPU kernel divides - the two differ by at most one ulp
    torch.testing.assert_close(se.sigmas.cpu(), so.sigmas, rtol=2.4e-7, atol=0)
    got = lp.refine_loop(m, se, lat0.clone().to(cuda), pe.to(cuda), pm.to(cuda), 4, ts_e).cpu()
    assert torch.equal(got[:, :, :4], lat0[:, :, :4])                 # the condition latents are never stepped
    rel = ((got - want).norm() / want.norm()).item()
    assert rel < 5e-3, rel


def test_refine_prepare_matches_oracle(cuda):
    """Stage-1 clip -> upsample -> pad -> VAE encode -> normalise -> noise to t_thresh, with the first-frame condition."""
    from worldforge_b200 import longcat_pipeline as lp
    from worldforge_b200.vae import WfWanVAE
    vcfg = wan_vae.VaeConfig(dim=8)
    PV = wan_vae.init_params(vcfg, 2)
    vid = torch.randint(0, 256, (9, 16, 24, 3), generator=g(0), dtype=torch.uint8)
    img = torch.rand(1, 3, 32, 48, generator=g(1)) * 2 - 1
    want, ncl, added, nf = ols.refine_prepare(vid, img, adapters.OracleVAE(PV, vcfg), 32, 48, g(7), t_thresh=0.6, num_cond_frames=1,
                                              spatial_refine_only=True)
    vae = WfWanVAE(PV, cuda, dim=8)
    got, ncl2, added2, nf2 = lp.refine_prepare(vid, img, vae, 32, 48, g(7), t_thresh=0.6, num_cond_frames=1, spatial_refine_only=True,
                                               device=cuda)
    assert (ncl, added, nf) == (ncl2, added2, nf2) == (4, 12, 9)
    assert got.shape == want.shape and got.dtype == torch.float32
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 5e-3, rel            # tf32 convolutions of the engine's VAE vs the fp32 oracle (tests/test_vae_gpu.py bound)
