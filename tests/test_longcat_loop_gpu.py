"""LongCat guided i2v loop (IRR + FLF + DSG, CFG-zero, Euler): engine vs oracle."""
import pytest
import torch

from oracle import adapters, longcat_dit as old, longcat_sched as ols, wan_vae

pytestmark = pytest.mark.gpu
KW = dict(hidden_size=128, depth=1, num_heads=1, caption_channels=32, adaln_tembed_dim=32, frequency_embedding_size=32)
KNOBS = dict(guidance_scale=4.0, guided=True, resample_steps=2, guide_steps=6, resample_round=7, omega=4.0, omega_resample=2.0,
             use_pca_channel_selection=True, max_replace_threshold=3)


def _inputs():
    from worldforge_b200 import synth
    inp = synth.make_inputs(9, 64, 96, text_len=8, text_dim=32, img_len=3, img_dim=16)
    g0 = torch.Generator().manual_seed(5)
    pe = torch.randn(2, 1, 8, 32, generator=g0).to(torch.bfloat16)
    pm = torch.ones(2, 8, dtype=torch.int64); pm[0, 5:] = 0
    return inp, pe, pm


@pytest.mark.parametrize("distill", [False, True])
def test_loop_kernels_match_oracle_with_shared_models(cuda, distill):
    """Same (oracle, CPU) DiT and VAE in both loops: only the scheduler / CFG-zero / DSG / FLF kernels differ."""
    from worldforge_b200 import longcat_pipeline as wlp
    cfg, vcfg = old.LongCatConfig(**KW), wan_vae.VaeConfig(dim=8)
    P, PV = old.init_params(cfg, 3), wan_vae.init_params(vcfg, 2)
    inp, pe, pm = _inputs()

    def run(loop, sched):
        hist = []
        loop(adapters.OracleLongCatDit(P, cfg, amp=True), adapters.OracleVAE(PV, vcfg), sched, inp.latents.clone().to(cuda),
             pe.to(cuda), pm.to(cuda), 8, use_distill=distill, video_ref=inp.video_ref.to(cuda), mask=inp.mask.to(cuda),
             generator=torch.Generator().manual_seed(42), on_step=lambda i, l: hist.append(l.detach().clone().cpu()), **KNOBS)
        return hist

    o, w = ols.OracleEuler(shift=1.0), wlp.WfFlowMatchEulerScheduler(shift=1.0)
    want, got = run(ols.denoise_loop, o), run(wlp.denoise_loop, w)
    assert o.flf_log == w.flf_log and o.fuse_calls == w.fuse_calls == 6
    assert any(len(c) >= 1 for _, c in w.flf_log)
    for a, b in zip(want, got):
        assert a.dtype == b.dtype == torch.float32
        # fp32 everywhere; the only freedom is the summation order of the CFG-zero / DSG reductions - an fp32-ulp change of the
        # latents that the (bf16) DiT of the next forward turns into occasional bf16 rounding flips
        assert ((a - b).norm() / a.norm()).item() < 2e-3


def test_longcat_end_to_end(cuda):
    """CUDA LongCat DiT + CUDA VAE + kernel scheduler against the all-oracle run, distilled 4-step schedule."""
    from worldforge_b200 import longcat, longcat_pipeline as wlp, vae as wvae
    kw = dict(hidden_size=256, depth=2, num_heads=2, caption_channels=32, adaln_tembed_dim=32, frequency_embedding_size=32)
    cfg, vcfg = old.LongCatConfig(**kw), wan_vae.VaeConfig(dim=8)
    P, PV = old.init_params(cfg, 3), wan_vae.init_params(vcfg, 2)
    inp, pe, pm = _inputs()
    knobs = dict(KNOBS, guide_steps=3, resample_round=3)
    want = []
    ols.denoise_loop(adapters.OracleLongCatDit(P, cfg, amp=True), adapters.OracleVAE(PV, vcfg), ols.OracleEuler(shift=1.0),
                     inp.latents.clone().to(cuda), pe.to(cuda), pm.to(cuda), 4, use_distill=True, video_ref=inp.video_ref.to(cuda),
                     mask=inp.mask.to(cuda), generator=torch.Generator().manual_seed(42),
                     on_step=lambda i, l: want.append(l.clone().cpu()), **knobs)
    dit = longcat.WfLongCatTransformer.from_state_dict(P, longcat.LongCatConfig(**kw), cuda)
    vae = wvae.WfWanVAE(PV, cuda, dim=8)
    got = []
    wlp.denoise_loop(dit, vae, wlp.WfFlowMatchEulerScheduler(shift=1.0), inp.latents.clone().to(cuda), pe.to(cuda), pm.to(cuda), 4,
                     use_distill=True, video_ref=inp.video_ref.to(cuda), mask=inp.mask.to(cuda),
                     generator=torch.Generator().manual_seed(42), on_step=lambda i, l: got.append(l.clone().cpu()), **knobs)
    assert dit.calls == 2 * (3 * 2 + 1)
    for i, (a, b) in enumerate(zip(want, got)):
        rel = ((a - b).norm() / a.norm()).item()
        assert rel < 2e-2, (i, rel)     # bf16 residual stream in the DiT (5e-3 by itself) + tf32 VAE round trip
