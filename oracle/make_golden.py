"""Generate tests/golden/wan_golden.pt from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  The reference pins nothing itself (no tests, no golden vectors), so the
fixtures are outputs of the reference's own modules, imported from /root/reference through oracle/ref_shim.py, on
inputs and random-init weights that are reproducible from seeds (oracle.*.init_params, worldforge_b200.synth):

  dit_fp32    WanModel.forward in fp32                                   (wan/modules/model.py:493-582)
  dit_amp     the same under bf16 autocast with the module's fp32 islands (the GPU dtype flow, run on the CPU)
  vae_mu/dec  WanVAE_.encode / .decode, chunked with the feature cache   (wan/modules/vae.py:516-568)
  sched_*     14 guided steps (IRR + FLF + DSG) driven through the reference's UniPCMultistepScheduler
              (utils/scheduling_unipc_multistep_clean.py), per-step latents and the FLF channel choices
  wan_pipeline_call   WanImageToVideoPipeline.__call__ itself (utils/pipeline_wan_i2v_clean.py:390-753), unmodified, over
              the oracle DiT / VAE and the reference scheduler: conditioning tensor and per-step latents

Run:  python -m oracle.make_golden     (the GPU box never runs this; it only reads the committed file)
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("TRITON_INTERPRET", "1")      # the reference's Triton BSA kernels run under Triton's CPU interpreter
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")   # ... and its @torch.compile gating helpers eagerly

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import adapters, longcat_bsa, longcat_dit, longcat_sched, pipeline, ref_shim, wan_dit, wan_vae  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "wan_golden.pt")

DIT_CFG = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16, img_dim=32, img_len=5, freq_dim=32)
DIT_GRID = (3, 8, 12)
SCHED_DIT = dict(dim=128, ffn_dim=256, num_heads=1, num_layers=1, text_dim=32, text_len=8, img_dim=16, img_len=3, freq_dim=32)
SCHED_KNOBS = dict(guided=True, resample_steps=2, guide_steps=12, omega=4.0, omega_resample=4.0, resample_round=13,
                   use_pca_channel_selection=True, static=True)


def dit_inputs():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(36, *DIT_GRID, generator=g)
    ctx = torch.randn(16, 64, generator=g)
    clip = torch.randn(5, 32, generator=g)
    return x, torch.tensor([737]), ctx, clip


def ref_dit(amp: bool):
    cfg = wan_dit.DitConfig(**DIT_CFG)
    P = wan_dit.init_params(cfg, 7)
    mod = ref_shim.load_wan_model_module(cpu_autocast=amp)
    mod.T5_CONTEXT_TOKEN_NUMBER = cfg.text_len      # the I2V cross-attention splits image / text tokens at this constant
    m = mod.WanModel(model_type="i2v", in_dim=36, dim=cfg.dim, ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim,
                     text_dim=cfg.text_dim, out_dim=16, num_heads=cfg.num_heads, num_layers=cfg.num_layers, text_len=cfg.text_len)
    m.img_emb = mod.MLPProj(cfg.img_dim, cfg.dim)     # the image-embedding width is hard-coded to 1280 in the constructor
    m.load_state_dict({k: v.clone() for k, v in P.items()})
    m.eval()
    x, t, ctx, clip = dit_inputs()
    L = DIT_GRID[0] * DIT_GRID[1] * DIT_GRID[2] // 4
    with torch.no_grad():
        if amp:
            bf = torch.bfloat16
            with torch.autocast("cpu", dtype=bf):
                return m([x[:16].to(bf)], t, [ctx.to(bf)], seq_len=L, clip_fea=clip.unsqueeze(0).to(bf), y=[x[16:].to(bf)])[0]
        return m([x[:16]], t, [ctx], seq_len=L, clip_fea=clip.unsqueeze(0), y=[x[16:]])[0]


def ref_vae():
    cfg = wan_vae.VaeConfig(dim=16)
    P = wan_vae.init_params(cfg, 5)
    vm = ref_shim.load_wan_vae_module()
    m = vm.WanVAE_(dim=16, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                   temperal_downsample=[False, True, True], dropout=0.0).eval()
    sd = m.state_dict()
    m.load_state_dict({k: P[k].reshape(sd[k].shape) for k in sd})
    g = torch.Generator().manual_seed(1)
    video = torch.rand(3, 9, 32, 48, generator=g) * 2 - 1
    z = torch.randn(16, 3, 4, 6, generator=g)
    with torch.no_grad():
        mu = m.encode(video.unsqueeze(0), [0.0, 1.0])[0]
        dec = m.decode(z.unsqueeze(0), [0.0, 1.0])[0].clamp(-1, 1)
    return mu, dec


def sched_setup():
    from worldforge_b200 import synth
    dcfg = wan_dit.DitConfig(**SCHED_DIT)
    vcfg = wan_vae.VaeConfig(dim=8)
    PD, PV = wan_dit.init_params(dcfg, 1), wan_vae.init_params(vcfg, 2)
    inp = synth.make_inputs(9, 64, 96, text_len=8, text_dim=32, img_len=3, img_dim=16)
    return dcfg, vcfg, PD, PV, inp


def run_sched(sched, steps=14):
    dcfg, vcfg, PD, PV, inp = sched_setup()
    hist = []
    pipeline.denoise_loop(adapters.OracleTransformer(PD, dcfg, amp=True), adapters.OracleVAE(PV, vcfg), sched,
                          inp.latents.clone(), inp.condition, inp.prompt_embeds, inp.negative_prompt_embeds, inp.image_embeds,
                          steps, 4.0, video_ref=inp.video_ref, mask=inp.mask, generator=torch.Generator().manual_seed(42),
                          on_step=lambda i, l: hist.append(l.clone()), **SCHED_KNOBS)
    return hist


def ref_sched():
    sm = ref_shim.load_scheduler_module()
    s = sm.UniPCMultistepScheduler(num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction",
                                   use_flow_sigmas=True, flow_shift=3.0)
    # record the FLF choices the reference's selector makes
    log = []
    orig = sm.VideoMotionPCASelector.select_motion_related_channels
    def spy(self, *a, **k):
        r = orig(self, *a, **k)
        log.append((k.get("current_step"), list(r)))
        return r
    sm.VideoMotionPCASelector.select_motion_related_channels = spy
    with ref_shim.no_silent_fallbacks(sm.VideoMotionPCASelector) as events:
        hist = run_sched(s)
    sm.VideoMotionPCASelector.select_motion_related_channels = orig
    assert not events, f"the reference took a silent fallback: {events}"
    s50 = sm.UniPCMultistepScheduler(num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction",
                                     use_flow_sigmas=True, flow_shift=3.0)
    s50.set_timesteps(50)
    return hist, log, dict(timesteps=s50.timesteps.clone(), sigmas=s50.sigmas.clone(),
                           resample_timesteps=s50.resample_timesteps.clone())


PIPE_STEPS = 14
# steps 0-11 guided (FLF selects 0 / 1 / 2-6 channels as the step index passes 5 and 10); omega != omega_resample and
# resample_round > guide_steps: step 12 runs IRR + DSG without FLF and switches omega for good (:678-679); step 13 is plain
PIPE_KNOBS = dict(guided=True, resample_steps=2, guide_steps=12, omega=4.0, omega_resample=2.0, resample_round=13,
                  use_pca_channel_selection=True, static=True)


def pipeline_inputs():
    dcfg, vcfg, PD, PV, inp = sched_setup()
    image = inp.video_ref[:, :, 0] * 2 - 1                      # [1,3,H,W] in [-1,1]: the first reference frame
    return dcfg, vcfg, PD, PV, inp, image


def run_pipeline_oracle(sched):
    """oracle/pipeline.py (prepare_condition + denoise_loop) on the inputs of ref_pipeline_call."""
    dcfg, vcfg, PD, PV, inp, image = pipeline_inputs()
    vae = adapters.OracleVAE(PV, vcfg)
    cond = pipeline.prepare_condition(vae, image, 9, 64, 96)
    hist = []
    pipeline.denoise_loop(adapters.OracleTransformer(PD, dcfg, amp=True), vae, sched, inp.latents.clone().float(), cond,
                          inp.prompt_embeds.to(torch.bfloat16), inp.negative_prompt_embeds.to(torch.bfloat16),
                          inp.image_embeds.to(torch.bfloat16), PIPE_STEPS, 4.0, video_ref=inp.video_ref, mask=inp.mask,
                          generator=torch.Generator().manual_seed(42), on_step=lambda i, l: hist.append(l.clone()), **PIPE_KNOBS)
    return cond, hist


def ref_pipeline_call():
    """WanImageToVideoPipeline.__call__ of the reference (utils/pipeline_wan_i2v_clean.py:390-753), UNMODIFIED, driving the
    reference's own UniPCMultistepScheduler over the oracle DiT / VAE objects: prepare_latents, the IRR inner loop, CFG,
    re-noise, DSG with its scheduler-state pokes and the bf16 cast.  Per-step latents through callback_on_step_end."""
    pm, sm = ref_shim.load_wan_pipeline_module(), ref_shim.load_scheduler_module()
    dcfg, vcfg, PD, PV, inp, image = pipeline_inputs()
    enc = ref_shim.FixedImageEncoder(inp.image_embeds)
    sched = sm.UniPCMultistepScheduler(num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction",
                                       use_flow_sigmas=True, flow_shift=3.0)
    pipe = pm.WanImageToVideoPipeline(tokenizer=None, text_encoder=None, image_encoder=enc.encoder, image_processor=enc.processor,
                                      transformer=adapters.OracleTransformer(PD, dcfg, amp=True), vae=adapters.OracleVAE(PV, vcfg),
                                      scheduler=sched)
    hist, conds = [], []
    orig_prepare = pipe.prepare_latents
    def spy_prepare(*a, **k):                                   # record the conditioning tensor the reference builds
        lat, cond = orig_prepare(*a, **k)
        conds.append(cond.clone())
        return lat, cond
    pipe.prepare_latents = spy_prepare
    def on_step(p, i, t, kw):
        hist.append(kw["latents"].clone())
        return {}
    with ref_shim.no_silent_fallbacks(sm.VideoMotionPCASelector) as events:
        out = pipe(image=image, height=64, width=96, num_frames=9, num_inference_steps=PIPE_STEPS, guidance_scale=4.0,
                   generator=torch.Generator().manual_seed(42), latents=inp.latents.clone(), prompt_embeds=inp.prompt_embeds,
                   negative_prompt_embeds=inp.negative_prompt_embeds, output_type="latent", return_dict=False,
                   callback_on_step_end=on_step, video_ref=inp.video_ref, mask=inp.mask, **PIPE_KNOBS)[0]
    assert not events, f"the reference took a silent fallback: {events}"
    assert torch.equal(out, hist[-1])
    return dict(condition=conds[0], latents=[h.clone() for h in hist], dtypes=[str(h.dtype) for h in hist])


LC_DIT = dict(hidden_size=256, depth=2, num_heads=2, caption_channels=64, adaln_tembed_dim=32, frequency_embedding_size=32)
LC_SCHED_DIT = dict(hidden_size=128, depth=1, num_heads=1, caption_channels=32, adaln_tembed_dim=32, frequency_embedding_size=32)
LC_KNOBS = dict(guidance_scale=4.0, guided=True, resample_steps=2, guide_steps=6, resample_round=7, omega=4.0, omega_resample=2.0,
                use_pca_channel_selection=True, max_replace_threshold=3)


def longcat_dit_inputs():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 16, 3, 8, 12, generator=g)
    ts = torch.tensor([[0.0, 750.0, 750.0]])
    ctx = torch.randn(1, 1, 10, 64, generator=g)
    mask = torch.ones(1, 10, dtype=torch.int64); mask[:, 7:] = 0
    return x, ts, ctx, mask


def ref_longcat_dit():
    """LongCatVideoTransformer3DModel.forward in fp32 (longcat_video/modules/longcat_video_dit.py:280-370), i2v form."""
    cfg = longcat_dit.LongCatConfig(**LC_DIT)
    P = longcat_dit.init_params(cfg, 3)
    mod = ref_shim.load_longcat_dit_module()
    m = mod.LongCatVideoTransformer3DModel(in_channels=16, out_channels=16, hidden_size=256, depth=2, num_heads=2, caption_channels=64,
                                           mlp_ratio=4, adaln_tembed_dim=32, frequency_embedding_size=32, enable_xformers=True,
                                           cp_split_hw=[1, 1]).eval()
    m.load_state_dict({k: P[k].clone() for k in m.state_dict()})
    x, ts, ctx, mask = longcat_dit_inputs()
    with torch.no_grad():
        return m(x, ts, ctx, encoder_attention_mask=mask, num_cond_latents=1)[0]


def longcat_sched_inputs():
    from worldforge_b200 import synth
    inp = synth.make_inputs(9, 64, 96, text_len=8, text_dim=32, img_len=3, img_dim=16)
    g0 = torch.Generator().manual_seed(5)
    pe = torch.randn(2, 1, 8, 32, generator=g0).to(torch.bfloat16)
    pm = torch.ones(2, 8, dtype=torch.int64); pm[0, 5:] = 0
    return inp, pe, pm


def run_longcat_sched(sched, distill: bool):
    cfg, vcfg = longcat_dit.LongCatConfig(**LC_SCHED_DIT), wan_vae.VaeConfig(dim=8)
    P, PV = longcat_dit.init_params(cfg, 3), wan_vae.init_params(vcfg, 2)
    inp, pe, pm = longcat_sched_inputs()
    hist = []
    longcat_sched.denoise_loop(adapters.OracleLongCatDit(P, cfg, amp=True), adapters.OracleVAE(PV, vcfg), sched, inp.latents.clone(),
                               pe, pm, 8, use_distill=distill, video_ref=inp.video_ref, mask=inp.mask,
                               generator=torch.Generator().manual_seed(42), on_step=lambda i, l: hist.append(l.clone()), **LC_KNOBS)
    return hist


def ref_longcat_sched():
    """8 guided steps through the reference's FlowMatchEulerDiscreteScheduler (standard and distilled schedules)."""
    sm = ref_shim.load_longcat_scheduler_module()
    out = {}
    for distill in (False, True):
        log = []
        orig = sm.VideoMotionChannelSelector.select_motion_related_channels
        def spy(self, *a, _orig=orig, **k):
            r = _orig(self, *a, **k)
            log.append((k.get("current_step"), list(r)))
            return r
        sm.VideoMotionChannelSelector.select_motion_related_channels = spy
        with ref_shim.no_silent_fallbacks(sm.VideoMotionChannelSelector) as events:
            hist = run_longcat_sched(sm.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0), distill)
        sm.VideoMotionChannelSelector.select_motion_related_channels = orig
        assert not events, f"the reference took a silent fallback: {events}"
        out["distill" if distill else "standard"] = dict(latents=torch.stack(hist), flf=log)
    return out


BSA_CASES = {      # name: (grid_q, grid_k, heads, dtype, kwargs of flash_attn_bsa_3d)
    "c64_topk_f32": ((4, 8, 8), (4, 8, 8), 1, torch.float32, dict(sparsity=0.5, chunk_3d_shape_q=[4, 4, 4], chunk_3d_shape_k=[4, 4, 4])),
    "c64_cross_f16": ((4, 4, 8), (8, 4, 8), 2, torch.float16, dict(sparsity=0.5, chunk_3d_shape_q=[4, 4, 4], chunk_3d_shape_k=[4, 4, 4])),
    "c128_topk_f16": ((8, 8, 8), (8, 8, 8), 2, torch.float16, dict(sparsity=0.5, chunk_3d_shape_q=[4, 4, 8], chunk_3d_shape_k=[4, 4, 8])),
    "c64_cdf_topk_f16": ((4, 8, 8), (4, 8, 8), 2, torch.float16, dict(sparsity=0.75, cdf_threshold=0.6, chunk_3d_shape_q=[4, 4, 4],
                                                                      chunk_3d_shape_k=[4, 4, 4])),
}


def bsa_inputs(name):
    gq, gk, heads, dt, kw = BSA_CASES[name]
    g = torch.Generator().manual_seed(len(name))
    mk = lambda grid: torch.randn(1, heads, grid[0] * grid[1] * grid[2], 128, generator=g).to(dt)
    return mk(gq), mk(gk), mk(gk), gq, gk, kw


def ref_bsa():
    """flash_attn_bsa_3d of the reference (bsa_interface.py:612-659), Triton kernels through Triton's interpreter."""
    bsa = ref_shim.load_longcat_bsa_module()
    out = {}
    for name in BSA_CASES:
        q, k, v, gq, gk, kw = bsa_inputs(name)
        out[name] = bsa.flash_attn_bsa_3d(q, k, v, gq, gk, **kw)[0].clone()
    return out


LC_BSA = dict(sparsity=0.5, cdf_threshold=None, chunk_3d_shape_q=[4, 4, 4], chunk_3d_shape_k=[4, 4, 4])


def longcat_bsa_dit_inputs():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 16, 8, 16, 16, generator=g)
    ts = torch.tensor([[0.0] * 4 + [600.0] * 4])
    ctx = torch.randn(1, 1, 10, 64, generator=g)
    mask = torch.ones(1, 10, dtype=torch.int64); mask[:, 8:] = 0
    return x, ts, ctx, mask


def ref_longcat_dit_bsa():
    """The reference DiT with enable_bsa() (longcat_video_dit.py:272-274), fp32, 4 condition + 4 noise latent frames."""
    cfg = longcat_dit.LongCatConfig(**LC_DIT)
    P = longcat_dit.init_params(cfg, 3)
    mod = ref_shim.load_longcat_dit_module(bsa=True)
    m = mod.LongCatVideoTransformer3DModel(in_channels=16, out_channels=16, hidden_size=256, depth=2, num_heads=2, caption_channels=64,
                                           mlp_ratio=4, adaln_tembed_dim=32, frequency_embedding_size=32, enable_xformers=True,
                                           bsa_params=dict(LC_BSA), cp_split_hw=[1, 1]).eval()
    m.load_state_dict({k: P[k].clone() for k in m.state_dict()})
    m.enable_bsa()
    x, ts, ctx, mask = longcat_bsa_dit_inputs()
    with torch.no_grad():
        return m(x, ts, ctx, encoder_attention_mask=mask, num_cond_latents=4)[0]


def run_refine(sched):
    """The refine loop (6 of 10 steps, t_thresh 0.6) on the oracle DiT with BSA, driven through ``sched``."""
    cfg = longcat_dit.LongCatConfig(**LC_SCHED_DIT)
    P = longcat_dit.init_params(cfg, 3)
    g = torch.Generator().manual_seed(21)
    lat = torch.randn(1, 16, 8, 16, 16, generator=g)
    pe = torch.randn(1, 1, 8, 32, generator=g).to(torch.bfloat16)
    pm = torch.ones(1, 8, dtype=torch.int64); pm[0, 6:] = 0
    ts = longcat_sched.refine_schedule(sched, 10, 0.6)
    dit = adapters.OracleLongCatDit(P, cfg, amp=True, bsa=dict(LC_BSA))
    return longcat_sched.refine_loop(dit, sched, lat, pe, pm, 4, ts), ts.clone(), sched.sigmas.clone()


def ref_refine():
    sm = ref_shim.load_longcat_scheduler_module()
    lat, ts, sig = run_refine(sm.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0))
    return dict(latents=lat.clone(), timesteps=ts, sigmas=sig)


def main():
    assert ref_shim.available(), "/root/reference is not mounted"
    torch.manual_seed(0)
    mu, dec = ref_vae()
    hist, flf_log, tables = ref_sched()
    gold = {
        "dit_fp32": ref_dit(False).clone(), "dit_amp": ref_dit(True).float().clone(),
        "vae_mu": mu.clone(), "vae_dec": dec.clone(),
        "sched_latents": torch.stack([h.float() for h in hist]), "sched_dtype": str(hist[-1].dtype), "sched_flf": flf_log,
        "sched_tables_50": tables,
        "longcat_dit_fp32": ref_longcat_dit().clone(), "longcat_sched": ref_longcat_sched(),
        "bsa": ref_bsa(), "longcat_dit_bsa_fp32": ref_longcat_dit_bsa().clone(), "longcat_refine": ref_refine(),
        "wan_pipeline_call": ref_pipeline_call(),
        "meta": {"reference_commit": "3314da5", "torch": torch.__version__, "generator": "oracle/make_golden.py"},
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
