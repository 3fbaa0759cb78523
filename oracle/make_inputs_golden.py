"""Writes tests/golden/inputs_golden.pt: outputs of the reference's own ``soften_mask`` (infer_worldforge.py:105-150, taken
from the script with ``ast``) on seeded masks.  Run where /root/reference is mounted:  python -m oracle.make_inputs_golden"""
import os

import numpy as np
import torch

from oracle import inputs as oin

REF = "/root/reference/wan_for_worldforge/infer_worldforge.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs_golden.pt")


def masks(seed: int = 0) -> np.ndarray:
    """uint8 [6, 70, 90]: blobs, a half plane, thin structures, an all-set and an all-unset frame, fractional edge values."""
    rng = np.random.default_rng(seed)
    F, H, W = 6, 70, 90
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((F, H, W), np.uint8)
    m[0] = (((yy - 30) ** 2 + (xx - 40) ** 2) < 20 ** 2) * 255
    m[1] = (xx > 37) * 255
    m[1, 10:12, :] = 0; m[1, :, 60] = 0
    m[2] = (rng.random((H, W)) > 0.02) * 255
    m[3] = 255
    m[4] = 0
    m[5] = (yy + xx > 60) * 255
    m[5][(yy + xx > 60) & (yy + xx < 64)] = 128                    # resized masks carry in-between values
    return m


def main():
    ref = oin.reference_soften_mask(REF)
    mu8 = masks()
    arr = oin.stack_masks(mu8)
    gold = {"masks_u8": torch.from_numpy(mu8)}
    for td, kind in ((15, "sine"), (7, "linear"), (10, "exponential"), (4, "cosine")):
        gold[f"soft_{td}_{kind}"] = torch.from_numpy(ref(arr, td, kind))
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__" and len(__import__("sys").argv) == 1:
    main()


# ---------------------------------------------------------------------------------------------------------------------
# LongCat video continuation (KV cache): python -m oracle.make_inputs_golden kv  ->  tests/golden/longcat_kv_golden.pt
# ---------------------------------------------------------------------------------------------------------------------
KV_OUT = os.path.join(os.path.dirname(OUT), "longcat_kv_golden.pt")


def kv_inputs():
    g = torch.Generator().manual_seed(4)
    lat = torch.randn(1, 16, 5, 8, 12, generator=g)               # 2 clean condition frames + 3 noise frames
    ctx = torch.randn(1, 1, 10, 64, generator=g)
    mask = torch.ones(1, 10, dtype=torch.int64); mask[:, 7:] = 0
    return lat, ctx, mask


def ref_kv():
    """The reference DiT's caching pass (return_kv, skip_crs_attn; pipeline_longcat_video.py:336-350) on the condition
    frames and its forward over the noise frames with that cache (:1202-1209), fp32."""
    from oracle import longcat_dit, make_golden as mg, ref_shim
    cfg = longcat_dit.LongCatConfig(**mg.LC_DIT)
    P = longcat_dit.init_params(cfg, 3)
    mod = ref_shim.load_longcat_dit_module()
    m = mod.LongCatVideoTransformer3DModel(in_channels=16, out_channels=16, hidden_size=256, depth=2, num_heads=2, caption_channels=64,
                                           mlp_ratio=4, adaln_tembed_dim=32, frequency_embedding_size=32, enable_xformers=True,
                                           cp_split_hw=[1, 1]).eval()
    m.load_state_dict({k: P[k].clone() for k in m.state_dict()})
    lat, ctx, mask = kv_inputs()
    with torch.no_grad():
        out_c, cache = m(lat[:, :, :2], torch.zeros(1, 2), torch.zeros(1, 1, 10, 64), return_kv=True, skip_crs_attn=True)
        out_n = m(lat[:, :, 2:], torch.full((1, 3), 600.0), ctx, encoder_attention_mask=mask, num_cond_latents=2, kv_cache_dict=cache)
        joint = m(lat, torch.tensor([[0.0, 0.0, 600.0, 600.0, 600.0]]), ctx, encoder_attention_mask=mask, num_cond_latents=2)
    return dict(cond_out=out_c[0].clone(), noise_out=out_n[0].clone(), joint_noise_out=joint[0][:, 2:].clone(),
                k0=cache[0][0][0].clone(), v1=cache[1][1][0].clone())


# ---------------------------------------------------------------------------------------------------------------------
# LongCat generate_i2v, the loop itself: python -m oracle.make_inputs_golden i2v -> tests/golden/longcat_i2v_golden.pt
# ---------------------------------------------------------------------------------------------------------------------
I2V_OUT = os.path.join(os.path.dirname(OUT), "longcat_i2v_golden.pt")
I2V_STEPS = 8


def ref_generate_i2v(distill: bool):
    """LongCatVideoPipeline.generate_i2v of the reference (pipeline_longcat_video.py:619-1006), UNMODIFIED, driving the
    reference's own FlowMatchEulerDiscreteScheduler over the oracle DiT / VAE objects: encode_prompt's masking and batching,
    prepare_latents, the IRR inner loop, CFG-zero, the sign flip, re-noise and DSG.  Two things are handed in rather than
    computed: the prompt embeddings (FixedTextEncoder) and the output size (the bucket lookup would pick 480 x 832; the run
    uses 64 x 96).  The VAE adapter's posterior sample is its mode.  Recorded: the prepared latents, the DiT's input latents
    at EVERY forward (every IRR round of every step) and the final latents."""
    from oracle import adapters, longcat_dit, make_golden as mg, ref_shim, wan_vae
    pm_mod = ref_shim.load_longcat_pipeline_module()
    sm = ref_shim.load_longcat_scheduler_module()
    cfg, vcfg = longcat_dit.LongCatConfig(**mg.LC_SCHED_DIT), wan_vae.VaeConfig(dim=8)
    P, PV = longcat_dit.init_params(cfg, 3), wan_vae.init_params(vcfg, 2)
    inp, pe, pmask = mg.longcat_sched_inputs()
    table = {"neg": (pe[0, 0], int(pmask[0].sum())), "pos": (pe[1, 0], int(pmask[1].sum()))}
    enc = ref_shim.FixedTextEncoder(table, pe.shape[2], pe.shape[3])
    vae = adapters.OracleVAE(PV, vcfg)
    vae.config.scale_factor_temporal, vae.config.scale_factor_spatial = 4, 8
    enc_fn = vae.encode
    def encode(x):
        o = enc_fn(x)
        o.latent_dist.sample = lambda generator=None: o.latent_dist.mode()
        return o
    vae.encode = encode
    dit = adapters.OracleLongCatDit(P, cfg, amp=True)
    seen = []
    call = dit.__call__
    class Rec:
        dtype, config, cp_split_hw = dit.dtype, dit.config, dit.cp_split_hw
        def __call__(self, hidden_states, **kw):
            seen.append(hidden_states[-1].clone())
            return call(hidden_states=hidden_states, **kw)
    sched = sm.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0)
    pipe = pm_mod.LongCatVideoPipeline(tokenizer=enc.tokenizer, text_encoder=enc, vae=vae, scheduler=sched, dit=Rec())
    pipe.device = "cpu"
    pipe.get_condition_shape = lambda image, resolution, scale_factor_spatial=32: (64, 96)
    prepared = []
    prep = pipe.prepare_latents
    def prepare_latents(**kw):
        out = prep(**kw)
        prepared.append(out.clone())
        return out
    pipe.prepare_latents = prepare_latents
    image = inp.video_ref[:, :, 0] * 2 - 1
    knobs = dict(mg.LC_KNOBS)
    with ref_shim.no_silent_fallbacks(sm.VideoMotionChannelSelector) as events:
        final = pipe.generate_i2v(image=image, prompt="pos", negative_prompt="neg", num_frames=9, num_inference_steps=I2V_STEPS,
                                  use_distill=distill, generator=torch.Generator().manual_seed(42), output_type="latent",
                                  max_sequence_length=pe.shape[2], video_ref=inp.video_ref, mask=inp.mask, **knobs)
    assert not events, f"the reference took a silent fallback: {events}"
    return dict(prepared=prepared[0], dit_inputs=torch.stack(seen), final=final.clone(), flf=[len(c) for _, c in getattr(sched, "flf_log", [])])


VC_OUT = os.path.join(os.path.dirname(OUT), "longcat_vc_golden.pt")
VC_STEPS = 12


def ref_generate_vc(use_kv_cache: bool):
    """LongCatVideoPipeline.generate_vc of the reference (pipeline_longcat_video.py:1010-1270), UNMODIFIED, over the oracle DiT /
    VAE and the reference's scheduler: 9 frames of which 5 condition (2 clean latent frames + 1 noise frame), CFG-zero,
    enhance_hf schedule (6 steps above t = 500 of the 12-step schedule + the 10-step uniform tail), with and without the KV
    cache.  Handed in as for generate_i2v: prompt embeddings, the 64 x 96 size; the VAE adapter's posterior sample is its mode
    and it evaluates in fp32 (the pipeline hands it a bf16 clip)."""
    from oracle import adapters, longcat_dit, make_golden as mg, ref_shim, wan_vae
    pm_mod = ref_shim.load_longcat_pipeline_module()
    sm = ref_shim.load_longcat_scheduler_module()
    cfg, vcfg = longcat_dit.LongCatConfig(**mg.LC_SCHED_DIT), wan_vae.VaeConfig(dim=8)
    P, PV = longcat_dit.init_params(cfg, 3), wan_vae.init_params(vcfg, 2)
    inp, pe, pmask = mg.longcat_sched_inputs()
    table = {"neg": (pe[0, 0], int(pmask[0].sum())), "pos": (pe[1, 0], int(pmask[1].sum()))}
    enc = ref_shim.FixedTextEncoder(table, pe.shape[2], pe.shape[3])
    vae = adapters.OracleVAE(PV, vcfg)
    vae.config.scale_factor_temporal, vae.config.scale_factor_spatial = 4, 8
    enc_fn = vae.encode
    def encode(x):
        o = enc_fn(x.float())
        o.latent_dist.sample = lambda generator=None: o.latent_dist.mode()
        return o
    vae.encode = encode
    dit = adapters.OracleLongCatDit(P, cfg, amp=True)
    seen = []
    call = dit.__call__
    class Rec:
        dtype, config, cp_split_hw = dit.dtype, dit.config, dit.cp_split_hw
        def __call__(self, hidden_states, **kw):
            seen.append(hidden_states[-1].clone())
            return call(hidden_states=hidden_states, **kw)
    sched = sm.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0)
    pipe = pm_mod.LongCatVideoPipeline(tokenizer=enc.tokenizer, text_encoder=enc, vae=vae, scheduler=sched, dit=Rec())
    pipe.device = "cpu"
    pipe.get_condition_shape = lambda video, resolution, scale_factor_spatial=32: (64, 96)
    prepared = []
    prep = pipe.prepare_latents
    def prepare_latents(**kw):
        out = prep(**kw)
        prepared.append(out.clone())
        return out
    pipe.prepare_latents = prepare_latents
    video = inp.video_ref * 2 - 1                                       # [1,3,9,64,96] in [-1,1]
    final = pipe.generate_vc(video=video, prompt="pos", negative_prompt="neg", num_frames=9, num_cond_frames=5,
                             num_inference_steps=VC_STEPS, guidance_scale=4.0, generator=torch.Generator().manual_seed(42),
                             output_type="latent", max_sequence_length=pe.shape[2], use_kv_cache=use_kv_cache, enhance_hf=True)
    return dict(prepared=prepared[0], dit_inputs=[t.clone() for t in seen], final=final.clone(), timesteps=sched.timesteps.clone(),
                sigmas=sched.sigmas.clone())


REFINE_OUT = os.path.join(os.path.dirname(OUT), "longcat_refine_call_golden.pt")


def refine_inputs():
    g = torch.Generator().manual_seed(31)
    stage1 = torch.randint(0, 256, (17, 40, 40, 3), generator=g, dtype=torch.uint8)        # 17 low-resolution frames
    image = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1                                 # first frame at the target size
    return stage1, image


def ref_generate_refine():
    """LongCatVideoPipeline.generate_refine of the reference (pipeline_longcat_video.py:1271-1511), UNMODIFIED, over the oracle
    DiT (block-sparse attention on) / VAE and the reference's scheduler: upsampling chain, BSA padding (1 condition frame ->
    4 condition latents, 16 noise frames -> 4 noise latents), VAE encode, noising to t = 0.6, Euler steps from there.  Handed in: the
    prompt embeddings and the 128 x 128 target size; the VAE adapter evaluates in fp32 and its posterior sample is its mode."""
    from oracle import adapters, longcat_dit, make_golden as mg, ref_shim, wan_vae
    pm_mod = ref_shim.load_longcat_pipeline_module()
    pm_mod.torch_gc = lambda: None                 # the module's helper empties the CUDA cache (:24-28); there is no CUDA here
    sm = ref_shim.load_longcat_scheduler_module()
    cfg, vcfg = longcat_dit.LongCatConfig(**mg.LC_SCHED_DIT), wan_vae.VaeConfig(dim=8)
    P, PV = longcat_dit.init_params(cfg, 3), wan_vae.init_params(vcfg, 2)
    _, pe, pmask = mg.longcat_sched_inputs()
    enc = ref_shim.FixedTextEncoder({"pos": (pe[1, 0], int(pmask[1].sum()))}, pe.shape[2], pe.shape[3])
    vae = adapters.OracleVAE(PV, vcfg)
    vae.config.scale_factor_temporal, vae.config.scale_factor_spatial = 4, 8
    enc_fn = vae.encode
    def encode(x):
        o = enc_fn(x.float())
        o.latent_dist.sample = lambda generator=None: o.latent_dist.mode()
        return o
    vae.encode = encode
    dit = adapters.OracleLongCatDit(P, cfg, amp=True, bsa=dict(mg.LC_BSA))
    seen = []
    call = dit.__call__
    class Rec:
        dtype, config, cp_split_hw = dit.dtype, dit.config, dit.cp_split_hw
        def __call__(self, hidden_states, **kw):
            seen.append(hidden_states[-1].clone())
            return call(hidden_states=hidden_states, **kw)
    sched = sm.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0)
    pipe = pm_mod.LongCatVideoPipeline(tokenizer=enc.tokenizer, text_encoder=enc, vae=vae, scheduler=sched, dit=Rec())
    pipe.device = "cpu"
    pipe.get_condition_shape = lambda video, resolution, scale_factor_spatial=32: (128, 128)
    prepared = []
    prep = pipe.prepare_latents
    def prepare_latents(**kw):
        out = prep(**kw)
        prepared.append(out.clone())
        return out
    pipe.prepare_latents = prepare_latents
    stage1, image = refine_inputs()
    final = pipe.generate_refine(image=image, prompt="pos", stage1_video=list(stage1.numpy()), num_cond_frames=1, num_inference_steps=10,
                                 generator=torch.Generator().manual_seed(42), output_type="latent", max_sequence_length=pe.shape[2],
                                 t_thresh=0.6, spatial_refine_only=True)
    return dict(prepared=prepared[0], dit_inputs=[t.clone() for t in seen], final=final.clone(), timesteps=sched.timesteps.clone())


if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "refine":
    torch.save(ref_generate_refine(), REFINE_OUT)
    print("wrote", REFINE_OUT, os.path.getsize(REFINE_OUT), "bytes")

if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "vc":
    torch.save({"kv": ref_generate_vc(True), "nokv": ref_generate_vc(False)}, VC_OUT)
    print("wrote", VC_OUT, os.path.getsize(VC_OUT), "bytes")

if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "i2v":
    torch.save({"standard": ref_generate_i2v(False), "distill": ref_generate_i2v(True)}, I2V_OUT)
    print("wrote", I2V_OUT, os.path.getsize(I2V_OUT), "bytes")

if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "kv":
    torch.save(ref_kv(), KV_OUT)
    print("wrote", KV_OUT, os.path.getsize(KV_OUT), "bytes")


# ---- fuse_latents with a mis-sized clip / mask (scheduling_unipc_multistep_clean.py:1297-1371) -----------------------
PRESIZE_OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "presize_golden.pt")


class RecordingVAE:
    """decode() hands back a fixed clip, encode() records the fused clip it is given: isolates the reference's resize +
    blend from the networks."""

    def __init__(self, decoded, z_dim=16):
        from types import SimpleNamespace
        from oracle import wan_vae
        self.decoded, self.fused = decoded, None
        self.config = SimpleNamespace(z_dim=z_dim, latents_mean=list(wan_vae.LATENTS_MEAN[:z_dim]), latents_std=list(wan_vae.LATENTS_STD[:z_dim]))

    def decode(self, z, return_dict=False):
        return (self.decoded.clone(),)

    def encode(self, x):
        from types import SimpleNamespace
        self.fused = x.clone()
        lat = torch.zeros(1, self.config.z_dim, (x.shape[2] - 1) // 4 + 1, x.shape[3] // 8, x.shape[4] // 8)
        return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: lat))


def presize_inputs():
    g = torch.Generator().manual_seed(77)
    dec = torch.rand(1, 3, 5, 32, 48, generator=g) * 2 - 1
    x0 = torch.randn(1, 16, 2, 4, 6, generator=g)
    cases = {
        "up": (torch.rand(1, 3, 5, 20, 30, generator=g), (torch.rand(1, 1, 5, 20, 30, generator=g) > 0.5).float()),
        "down_odd": (torch.rand(1, 3, 5, 45, 67, generator=g), torch.rand(1, 1, 5, 45, 67, generator=g)),
        "mask_3ch": (torch.rand(1, 3, 5, 32, 48, generator=g), torch.rand(1, 3, 5, 32, 48, generator=g)),
        "both_mask_3ch": (torch.rand(1, 3, 5, 24, 40, generator=g), torch.rand(1, 3, 5, 16, 24, generator=g)),
        "clip_only": (torch.rand(1, 3, 5, 64, 96, generator=g), torch.rand(1, 1, 5, 32, 48, generator=g)),
    }
    return dec, x0, cases


def ref_presize():
    """The reference scheduler's own fuse_latents on clips / masks that do not match the decoded clip: what it hands to
    vae.encode (the resized clip blended with the decoded one)."""
    from oracle import ref_shim
    sm = ref_shim.load_scheduler_module()
    s = sm.UniPCMultistepScheduler(num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction",
                                   use_flow_sigmas=True, flow_shift=3.0)
    dec, x0, cases = presize_inputs()
    out = {}
    for name, (clip, mask) in cases.items():
        vae = RecordingVAE(dec)
        s.fuse_latents(x0, clip, mask, vae=vae)
        out[name] = vae.fused
    return out


if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "presize":
    torch.save(ref_presize(), PRESIZE_OUT)
    print("wrote", PRESIZE_OUT, os.path.getsize(PRESIZE_OUT), "bytes")
