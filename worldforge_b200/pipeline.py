"""The WorldForge guided denoising loop (IRR + FLF + DSG) over the sm_100a engine.

Host code that stands where ``WanImageToVideoPipeline.__call__`` stands in the reference
(utils/pipeline_wan_i2v_clean.py:390-753): same keyword arguments, same control flow, same order
of random draws from the caller's CPU generator - every tensor expression of the loop is a fused
kernel launch:

  CFG   noise_pred + s*(noise_pred - noise_uncond)          (:611)      -> wf_cfg_combine
  IRR   re-noise of the fused x0 prediction                 (:642-660)  -> wf_renoise (scheduler.add_noise)
  DSG   three reductions + the guided combination           (:664-681)  -> wf_dsg
        followed by the manual UniP redo                    (:683-708)  -> wf_x0_convert, wf_unip_update

The text / image encoders run once per video and are outside the hot path:
the pipeline takes ``prompt_embeds``, ``negative_prompt_embeds`` and ``image_embeds`` like the
reference does when they are pre-computed (:403-405).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import lib


def prepare_condition(vae, image: torch.Tensor, num_frames: int, height: int, width: int) -> torch.Tensor:
    """The 20-channel I2V condition of prepare_latents (:327-362): 4 first-frame mask channels and the
    normalised VAE latents of [first frame, zeros...].  image: [1,3,H,W] fp32 in [-1,1] on the device."""
    from .synth import frame_mask_channels
    dev = image.device
    video = torch.cat([image.unsqueeze(2), image.new_zeros(1, 3, num_frames - 1, height, width)], dim=2)
    lat = vae.encode(video.to(torch.float32)).latent_dist.mode()
    z = vae.config.z_dim
    mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(dev, lat.dtype)
    inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(dev, lat.dtype)
    lat = (lat - mean) * inv_std
    mask = frame_mask_channels(num_frames, height // 8, width // 8).to(dev)
    return torch.cat([mask, lat], dim=1)


class GuidedSampler:
    """The loop of pipeline_wan_i2v_clean.py:556-728 as an object: ``begin()`` then ``step(i)`` for every
    timestep.  ``denoise_loop`` below is the plain loop; bench.py drives ``step`` itself so that it can place
    host<->device copies around each step."""

    def __init__(self, transformer, vae, scheduler, guidance_scale: float, guided: bool = False, resample_steps: int = 1,
                 guide_steps: int = 20, omega: float = 1.8, omega_resample: float = 1.0, resample_round: int = 20,
                 use_pca_channel_selection: bool = False, static: bool = False,
                 generator: Optional[torch.Generator] = None, cfg_parallel=None):
        self.transformer, self.vae, self.scheduler = transformer, vae, scheduler
        self.cfg_parallel = cfg_parallel       # worldforge_b200.ulysses.CfgParallel: this rank runs ONE of the two CFG forwards
        self.guidance_scale, self.guided, self.resample_steps = guidance_scale, guided, resample_steps
        self.guide_steps, self.omega, self.omega_resample, self.resample_round = guide_steps, omega, omega_resample, resample_round
        self.use_pca_channel_selection, self.static, self.generator = use_pca_channel_selection, static, generator
        self.forwards = 0

    def begin(self, num_inference_steps: int, device):
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        if not hasattr(self.scheduler, "derivative_history"):
            self.scheduler.derivative_history = []
        return self.scheduler.timesteps

    @torch.no_grad()
    def step(self, i: int, latents, condition, prompt_embeds, negative_prompt_embeds, image_embeds, video_ref=None,
             mask=None):
        sch, tr = self.scheduler, self.transformer
        device, tdtype = latents.device, tr.dtype
        t = sch.timesteps[i]
        do_cfg = self.guidance_scale > 1
        sch.derivative_history = []
        x0, out = None, None
        for r in range(self.resample_steps):
            if r > 0:
                sch.set_resample_mode(True)
                t_model = sch.get_resample_timestep(i).expand(latents.shape[0]).to(device=device)
                sch._step_index -= 1
                if sch.lower_order_nums > 0 and sch.last_lower_order_nums < sch.config.solver_order:
                    sch.lower_order_nums -= 1
                sch.this_order = sch.last_this_order
            else:
                sch.set_resample_mode(False)
                t_model = t.expand(latents.shape[0])
            model_in = torch.cat([latents, condition], dim=1).to(tdtype)
            if do_cfg and self.cfg_parallel is not None:
                mine = negative_prompt_embeds if self.cfg_parallel.branch else prompt_embeds
                with lib.phase("dit.forward"):
                    v = tr(hidden_states=model_in, timestep=t_model, encoder_hidden_states=mine,
                           encoder_hidden_states_image=image_embeds, attention_kwargs=None, return_dict=False)[0]
                self.forwards += 2                 # both forwards of the step ran, one in each half of the ranks
                with lib.phase("cfg.exchange"):
                    v, v_u = self.cfg_parallel.exchange(v)
                v = lib.cfg_combine(v.contiguous(), v_u.contiguous(), self.guidance_scale)
                if r < 1:
                    sch.derivative_history.append(v)
            else:
                with lib.phase("dit.forward"):
                    v = tr(hidden_states=model_in, timestep=t_model, encoder_hidden_states=prompt_embeds,
                           encoder_hidden_states_image=image_embeds, attention_kwargs=None, return_dict=False)[0]
                self.forwards += 1
            if do_cfg and self.cfg_parallel is None:
                with lib.phase("dit.forward"):
                    v_u = tr(hidden_states=model_in, timestep=t_model, encoder_hidden_states=negative_prompt_embeds,
                             encoder_hidden_states_image=image_embeds, attention_kwargs=None, return_dict=False)[0]
                self.forwards += 1
                v = lib.cfg_combine(v.contiguous(), v_u.contiguous(), self.guidance_scale)
                if r < 1:
                    sch.derivative_history.append(v)
            with lib.phase("scheduler.step"):
                out = sch.step(v, t, latents, mask=mask, guided=self.guided and i < self.guide_steps and r < self.resample_steps,
                               video_latents=video_ref, vae=self.vae, resampling=r > 0, return_dict=True, current_step=i,
                               resample_count=self.resample_steps, is_resample_round=i < self.resample_round,
                               use_pca_channel_selection=self.use_pca_channel_selection, static=self.static)
            if hasattr(out, "pred_x0"):
                x0 = out.pred_x0
            if i >= self.resample_round:
                break
            if r < self.resample_steps - 1 and x0 is not None:
                if self.generator is not None:   # CPU generator: the noise stream is part of the result (:643-645)
                    with lib.phase("irr.noise_draw_upload"):
                        noise = torch.randn(x0.shape, generator=self.generator).pin_memory().to(device=device, non_blocking=True)
                else:
                    noise = torch.randn(x0.shape, device=device)
                t_noise = sch.get_resample_timestep(i)
                if t_noise.dim() == 0:
                    t_noise = t_noise.unsqueeze(0)
                latents = sch.add_noise(x0, noise, t_noise.to(device=device), r, use_resample_sigma=True)

        if len(sch.derivative_history) > 1:                 # DSG (:664-708)
            g, w = sch.derivative_history[-1], sch.derivative_history[0]
            if i >= self.guide_steps:
                self.omega = self.omega_resample            # sticks for the rest of the run (:678-679)
            better = lib.dsg(g.contiguous(), w.contiguous(), self.omega)
            sch._step_index -= 1
            if sch.lower_order_nums > 0 and sch.last_lower_order_nums < sch.config.solver_order:
                sch.lower_order_nums -= 1
            m = sch.convert_model_output(better, sample=latents)
            sch.last_sample = latents
            sch.model_outputs[-1] = m
            latents = sch.multistep_uni_p_bh_update(model_output=better, sample=latents, order=sch.this_order)
            sch._step_index += 1
            if 0 <= sch.lower_order_nums < sch.config.solver_order:
                sch.lower_order_nums += 1
            latents = latents.to(dtype=tdtype)
        else:
            latents = out.prev_sample
        sch.set_resample_mode(False)
        return latents


@torch.no_grad()
def denoise_loop(transformer, vae, scheduler, latents, condition, prompt_embeds, negative_prompt_embeds,
                 image_embeds, num_inference_steps: int, guidance_scale: float, video_ref=None, mask=None,
                 guided: bool = False, resample_steps: int = 1, guide_steps: int = 20, omega: float = 1.8,
                 omega_resample: float = 1.0, resample_round: int = 20, use_pca_channel_selection: bool = False,
                 static: bool = False, generator: Optional[torch.Generator] = None,
                 on_step: Optional[Callable] = None, max_steps: Optional[int] = None, cfg_parallel=None) -> torch.Tensor:
    """``latents`` [1,16,f,h,w] fp32 on the device -> latents after the last step."""
    device = latents.device
    sampler = GuidedSampler(transformer, vae, scheduler, guidance_scale, guided, resample_steps, guide_steps, omega,
                            omega_resample, resample_round, use_pca_channel_selection, static, generator, cfg_parallel)
    timesteps = sampler.begin(num_inference_steps, device)
    if video_ref is not None and guided:
        video_ref = video_ref.to(device=device, dtype=torch.float32)
    if mask is not None and guided:
        mask = mask.to(device)
    for i in range(len(timesteps)):
        if max_steps is not None and i >= max_steps:
            break
        latents = sampler.step(i, latents, condition, prompt_embeds, negative_prompt_embeds, image_embeds, video_ref, mask)
        if on_step is not None:
            on_step(i, latents)
    return latents


class WfWanI2VPipeline:
    """``pipe(...)`` with the reference's keyword surface for pre-computed embeddings
    (infer_worldforge.py:290-309 -> pipeline_wan_i2v_clean.py:390-424)."""

    def __init__(self, transformer, vae, scheduler):
        self.transformer, self.vae, self.scheduler = transformer, vae, scheduler
        self.vae_scale_factor_temporal, self.vae_scale_factor_spatial = 4, 8
        self.cfg_parallel = None               # set to a worldforge_b200.ulysses.CfgParallel for the CFG x Ulysses layout

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def __call__(self, image=None, height: int = 480, width: int = 832, num_frames: int = 81,
                 num_inference_steps: int = 50, guidance_scale: float = 5.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, image_embeds=None, condition=None,
                 output_type: str = "latent", video_ref=None, mask=None, guided: bool = False, resample_steps: int = 1,
                 guide_steps: int = 20, omega: float = 1.8, omega_resample: float = 1.0, resample_round: int = 20,
                 use_pca_channel_selection: bool = False, static: bool = False, on_step=None, device="cuda"):
        if prompt_embeds is None or image_embeds is None:
            raise ValueError("pass prompt_embeds / negative_prompt_embeds / image_embeds: tokenisation and CLIP preprocessing stay with the "
                             "caller; worldforge_b200.encoders.t5_prompt_embeds / WfCLIPVisionEncoder compute the embeddings from token ids / pixels")
        if num_frames % 4 != 1:
            num_frames = max(num_frames // 4 * 4 + 1, 1)
        dev = torch.device(device)
        f, h, w = (num_frames - 1) // 4 + 1, height // 8, width // 8
        if latents is None:
            latents = torch.randn((1, self.vae.config.z_dim, f, h, w), generator=generator, dtype=torch.float32)
        latents = latents.to(device=dev, dtype=torch.float32)
        if condition is None:
            condition = prepare_condition(self.vae, image.to(dev, torch.float32), num_frames, height, width)
        td = self.transformer.dtype
        out = denoise_loop(self.transformer, self.vae, self.scheduler, latents, condition.to(dev),
                           prompt_embeds.to(dev, td), None if negative_prompt_embeds is None else negative_prompt_embeds.to(dev, td),
                           image_embeds.to(dev, td), num_inference_steps, guidance_scale, video_ref=video_ref, mask=mask,
                           guided=guided, resample_steps=resample_steps, guide_steps=guide_steps, omega=omega,
                           omega_resample=omega_resample, resample_round=resample_round,
                           use_pca_channel_selection=use_pca_channel_selection, static=static, generator=generator,
                           on_step=on_step, cfg_parallel=self.cfg_parallel)
        if output_type == "latent":
            return out
        from .scheduler import latent_stats
        mean_h, inv_std_h = latent_stats(self.vae.config.latents_mean, self.vae.config.latents_std, torch.float32)
        z = lib.latent_denorm(out.to(torch.float32).contiguous(), mean_h, inv_std_h)
        return self.vae.decode(z, return_dict=False)[0]
