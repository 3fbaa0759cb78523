"""Oracle: the WorldForge guided denoising loop (IRR + FLF + DSG), restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``WanImageToVideoPipeline.__call__`` in
``wan_for_worldforge/utils/pipeline_wan_i2v_clean.py`` :556-728 (outer loop
:563, IRR inner loop :573-662, CFG :602-611, scheduler.step :619-634, re-noise
:642-660, DSG :664-708) and ``prepare_latents`` :301-362.  The loop is written
against duck-typed ``transformer`` / ``vae`` / ``scheduler`` objects with the
surfaces listed in SURVEY.md §8(b), so the same function can drive the oracle
objects, the reference's own scheduler class, or the CUDA engine's objects.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch


def first_frame_mask(num_frames: int, lat_h: int, lat_w: int, t_scale: int = 4) -> torch.Tensor:
    """The 4-channel frame mask of prepare_latents (:353-360): [1, 4, f, h, w]."""
    m = torch.ones(1, 1, num_frames, lat_h, lat_w)
    m[:, :, 1:] = 0
    first = torch.repeat_interleave(m[:, :, 0:1], dim=2, repeats=t_scale)
    m = torch.cat([first, m[:, :, 1:]], dim=2)
    return m.view(1, -1, t_scale, lat_h, lat_w).transpose(1, 2)


def prepare_condition(vae, image: torch.Tensor, num_frames: int, height: int, width: int, t_scale: int = 4,
                      s_scale: int = 8) -> torch.Tensor:
    """The [B, 4 + z, f, h, w] conditioning tensor of prepare_latents (:316-362): the first frame followed by zeros is
    encoded (argmax = the mean), normalised with the VAE's latent statistics and preceded by the 4-channel frame mask.
    ``image`` [B,3,H,W] fp32 in [-1,1]."""
    video = torch.cat([image.unsqueeze(2), image.new_zeros(image.shape[0], image.shape[1], num_frames - 1, height, width)], dim=2)
    video = video.to(dtype=torch.float32)
    mean = torch.tensor(vae.config.latents_mean).view(1, vae.config.z_dim, 1, 1, 1).to(video.device, torch.float32)
    inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, vae.config.z_dim, 1, 1, 1).to(video.device, torch.float32)
    cond = (vae.encode(video).latent_dist.mode() - mean) * inv_std
    mask = first_frame_mask(num_frames, height // s_scale, width // s_scale, t_scale).to(cond.device)
    return torch.cat([mask.expand(cond.shape[0], -1, -1, -1, -1), cond], dim=1)


def denoise_loop(transformer, vae, scheduler, latents, condition, prompt_embeds, negative_prompt_embeds,
                 image_embeds, num_inference_steps: int, guidance_scale: float, video_ref=None, mask=None,
                 guided=False, resample_steps=1, guide_steps=20, omega=1.8, omega_resample=1.0,
                 resample_round=20, use_pca_channel_selection=False, static=False,
                 generator: Optional[torch.Generator] = None, transformer_dtype=torch.bfloat16,
                 on_step: Optional[Callable] = None, max_steps: Optional[int] = None) -> torch.Tensor:
    device = latents.device
    do_cfg = guidance_scale > 1
    scheduler.set_timesteps(num_inference_steps, device=device)
    timesteps = scheduler.timesteps
    if not hasattr(scheduler, "derivative_history"):
        scheduler.derivative_history = []
    out = None
    for i, t in enumerate(timesteps):
        if max_steps is not None and i >= max_steps:
            break
        scheduler.derivative_history = []
        x0 = None
        for r in range(resample_steps):
            if r > 0:
                scheduler.set_resample_mode(True)
                t_model = scheduler.get_resample_timestep(i).expand(latents.shape[0]).to(device=device)
                scheduler._step_index -= 1
                if scheduler.lower_order_nums > 0 and scheduler.last_lower_order_nums < scheduler.config.solver_order:
                    scheduler.lower_order_nums -= 1
                scheduler.this_order = scheduler.last_this_order
            else:
                scheduler.set_resample_mode(False)
                t_model = t.expand(latents.shape[0])
            model_in = torch.cat([latents, condition], dim=1).to(transformer_dtype)
            v = transformer(hidden_states=model_in, timestep=t_model, encoder_hidden_states=prompt_embeds,
                            encoder_hidden_states_image=image_embeds, attention_kwargs=None,
                            return_dict=False)[0]
            if do_cfg:
                v_u = transformer(hidden_states=model_in, timestep=t_model,
                                  encoder_hidden_states=negative_prompt_embeds,
                                  encoder_hidden_states_image=image_embeds, attention_kwargs=None,
                                  return_dict=False)[0]
                v = v + guidance_scale * (v - v_u)          # note: NOT v_u + s (v - v_u)
                if r < 1:
                    scheduler.derivative_history.append(v)
            out = scheduler.step(v, t, latents, mask=mask, guided=guided and i < guide_steps and r < resample_steps,
                                 video_latents=video_ref, vae=vae, resampling=r > 0, return_dict=True,
                                 current_step=i, resample_count=resample_steps, is_resample_round=i < resample_round,
                                 use_pca_channel_selection=use_pca_channel_selection, static=static)
            if hasattr(out, "pred_x0"):
                x0 = out.pred_x0
            if i >= resample_round:
                break
            if r < resample_steps - 1 and x0 is not None:
                if generator is not None:
                    noise = torch.randn(x0.shape, generator=generator).to(device=device)
                else:
                    noise = torch.randn(x0.shape, device=device)
                t_noise = scheduler.get_resample_timestep(i)
                if t_noise.dim() == 0:
                    t_noise = t_noise.unsqueeze(0)
                latents = scheduler.add_noise(x0, noise, t_noise.to(device=device), r, use_resample_sigma=True)

        if len(scheduler.derivative_history) > 1:           # DSG
            g, w = scheduler.derivative_history[-1], scheduler.derivative_history[0]
            dims = list(range(1, g.dim()))
            dot = torch.sum(g * w, dim=dims, keepdim=True)
            ng = torch.sqrt(torch.sum(g ** 2, dim=dims, keepdim=True))
            nw = torch.sqrt(torch.sum(w ** 2, dim=dims, keepdim=True))
            cos = dot / (ng * nw + 1e-8)
            sin = torch.sin(torch.acos(torch.clamp(cos, -1.0, 1.0)))
            ratio = ng / (nw + 1e-8)
            if i >= guide_steps:
                omega = omega_resample                      # sticks for the rest of the run (:678-679)
            better = g + omega * sin * (g - (ratio * cos) * w)

            scheduler._step_index -= 1
            if scheduler.lower_order_nums > 0 and scheduler.last_lower_order_nums < scheduler.config.solver_order:
                scheduler.lower_order_nums -= 1
            m = scheduler.convert_model_output(better, sample=latents)
            scheduler.last_sample = latents
            scheduler.model_outputs[-1] = m
            latents = scheduler.multistep_uni_p_bh_update(model_output=better, sample=latents,
                                                          order=scheduler.this_order)
            scheduler._step_index += 1
            if 0 <= scheduler.lower_order_nums < scheduler.config.solver_order:
                scheduler.lower_order_nums += 1
            latents = latents.to(dtype=transformer_dtype)
        else:
            latents = out.prev_sample
        scheduler.set_resample_mode(False)
        if on_step is not None:
            on_step(i, latents)
    return latents
