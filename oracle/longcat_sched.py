"""Oracle: LongCat's flow-matching Euler scheduler with FLF fusion and the guided i2v loop, restated on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``longcat_for_worldforge/longcat_video/modules/scheduling_flow_match_euler_discrete.py``
(set_timesteps :620-718, step :740-912, add_noise :1041-1070, fuse_latents :1072-1233, VideoMotionChannelSelector
:35-381) and ``longcat_video/pipeline_longcat_video.py`` (get_timesteps_sigmas :316-331, optimized_scale :374-383, the
denoising loop of generate_i2v :828-994).  Differences from the Wan path that matter: Euler update ``x + dt*v``; CFG-zero
(``st* = <v_c,v_u>/(|v_u|^2+1e-8)``) followed by a sign flip; per-frame timesteps with the clean first latent frame at
t = 0; FLF acts on the FULL latents (first frame included) and only when not resampling; its result feeds the re-noise
only - ``prev_sample`` ignores it (:900); per-CHANNEL min-max quantisation mapped onto the upper half of uint8
(:147-151); outliers use OR (:226); weights 0.4/0.4/0.2 (:237); selection policy :338-379.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from .flf import FARNEBACK


def _flows(video_u8: np.ndarray) -> torch.Tensor:
    import cv2
    fl = []
    for t in range(video_u8.shape[0] - 1):
        a = cv2.cvtColor(video_u8[t], cv2.COLOR_RGB2GRAY)
        b = cv2.cvtColor(video_u8[t + 1], cv2.COLOR_RGB2GRAY)
        fl.append(cv2.calcOpticalFlowFarneback(a, b, None, **FARNEBACK))
    return torch.from_numpy(np.stack(fl, 0).transpose(0, 3, 1, 2)).float().unsqueeze(0)


def quantise_channel(ch: torch.Tensor) -> np.ndarray:
    """[1,1,T,H,W] -> uint8 [T,H,W,3]: min-max over the channel, then (v+1)*127.5 (:329-330, :147-151)."""
    rgb = ch.repeat(1, 3, 1, 1, 1)
    rgb = (rgb - rgb.min()) / (rgb.max() - rgb.min() + 1e-8)
    v = rgb.to(torch.float32).squeeze(0).cpu().numpy().transpose(1, 2, 3, 0)
    if v.min() >= -1.1 and v.max() <= 1.1:
        return ((v + 1.0) * 127.5).clip(0, 255).astype(np.uint8)
    return (v * 255).clip(0, 255).astype(np.uint8)


def flow_similarity(ref: torch.Tensor, cand: torch.Tensor) -> float:
    d = ref - cand
    epe = torch.sqrt((d ** 2).sum(dim=2) + 1e-8)
    dot = (ref * cand).sum(dim=2)
    rn = torch.sqrt((ref ** 2).sum(dim=2) + 1e-8)
    cn = torch.sqrt((cand ** 2).sum(dim=2) + 1e-8)
    ang = torch.acos(torch.clamp(dot / (rn * cn + 1e-8), -1.0, 1.0)) * 180.0 / torch.pi
    outlier = (epe > 3.0) | (epe > rn * 0.05)
    err = (0.4 * torch.clamp(epe.mean() / 10.0, 0.0, 1.0) + 0.4 * torch.clamp(outlier.float().mean() / 0.5, 0.0, 1.0)
           + 0.2 * torch.clamp(ang.mean() / 30.0, 0.0, 1.0))
    return torch.clamp(1.0 - err, 0.0, 1.0).item()


def policy(scores, step: int, use_distill: bool, max_replace_threshold: Optional[int]) -> List[int]:
    s = np.array(scores)
    early = 3 if use_distill else 5
    if step <= early:
        out = np.argsort(s)[:1].tolist()
    else:
        cap = max_replace_threshold if max_replace_threshold is not None else (3 if use_distill else 1)
        thr = np.mean(s) - 0.625 * np.std(s)
        below = [i for i, v in enumerate(s) if v < thr]
        if len(below) < 1:
            out = np.argsort(s)[:1].tolist()
        elif len(below) > cap:
            out = [i for i, _ in sorted(((i, s[i]) for i in below), key=lambda p: p[1])[:cap]]
        else:
            out = below
    return sorted(out)


def select_channels(pred_x0, enc, step, use_distill=False, max_replace_threshold=None) -> List[int]:
    if step < 2:
        return []
    enc = enc.to(pred_x0.device, dtype=pred_x0.dtype)
    scores = []
    for c in range(pred_x0.shape[1]):
        pm = _flows(quantise_channel(pred_x0[:, c:c + 1]))
        rm = _flows(quantise_channel(enc[:, c:c + 1]))
        scores.append(flow_similarity(rm, pm))
    return policy(scores, step, use_distill, max_replace_threshold)


class StepOutput:
    def __init__(self, prev_sample, pred_x0=None):
        self.prev_sample, self.pred_x0 = prev_sample, pred_x0


class OracleEuler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift, stochastic_sampling=False)
        self.shift = shift
        self._step_index = None
        self._begin_index = None
        self.derivative_history = []
        self.is_resampling = False
        self.flf_log = []
        self.fuse_calls = 0

    step_index = property(lambda self: self._step_index)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None):
        sig = (sigmas.detach().cpu().numpy() if isinstance(sigmas, torch.Tensor) else np.asarray(sigmas)).astype(np.float32)
        sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        sig = torch.from_numpy(sig).to(dtype=torch.float32, device=device)
        self.timesteps = sig * self.config.num_train_timesteps
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device)])
        self.num_inference_steps = len(sig)
        self._step_index = None
        self._begin_index = None
        self.derivative_history = []

    def set_resample_mode(self, enabled: bool):
        self.is_resampling = enabled

    def _index_for_timestep(self, timestep, schedule=None):
        schedule = self.timesteps if schedule is None else schedule
        idx = (schedule == timestep).nonzero()
        return idx[1 if len(idx) > 1 else 0].item()

    def fuse_latents(self, pred_x0, video_ref, mask, vae, use_pca_channel_selection=False, current_step=0,
                     use_distill=False, max_replace_threshold=None, **kw):
        if mask is None or video_ref is None or vae is None:
            return pred_x0
        self.fuse_calls += 1
        dt, dev = pred_x0.dtype, pred_x0.device
        z = vae.config.z_dim
        mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(dev, dt)
        inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(dev, dt)
        dec = vae.decode((pred_x0 / inv_std + mean).to(dtype=vae.dtype), return_dict=False)[0]
        if video_ref.shape != dec.shape:
            raise ValueError("Dimension mismatch")
        ref = 2.0 * video_ref.to(dec.device, dec.dtype) - 1.0
        m = mask.to(dec.device, dec.dtype).repeat(1, dec.shape[1], 1, 1, 1)
        fused = (ref * m + dec * (1 - m)).to(dtype=vae.dtype)
        enc = vae.encode(fused).latent_dist.mode()
        enc = (enc - mean) * inv_std
        if use_pca_channel_selection:
            chans = select_channels(pred_x0, enc, current_step, use_distill, max_replace_threshold)
            self.flf_log.append((current_step, list(chans)))
            for c in chans:
                enc[:, c] = pred_x0[:, c]
        return enc.to(dev, dt)

    def step(self, model_output, timestep, sample, return_dict=True, video_ref=None, mask=None, guided=False,
             resampling=False, vae=None, use_pca_channel_selection=False, static=False, current_step=-1, total_steps=50,
             sample_full=None, use_distill=False, max_replace_threshold=None):
        if self._step_index is None:
            self._step_index = self._index_for_timestep(timestep.to(self.timesteps.device)) if self._begin_index is None else self._begin_index
        sample = sample.to(torch.float32)
        sigma, sigma_next = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        dt = sigma_next - sigma
        pred_x0 = sample - sigma * model_output
        if guided and video_ref is not None and not resampling and sample_full is not None:
            full = sample_full - sigma * torch.cat([torch.zeros_like(model_output[:, :, 0:1]), model_output], dim=2)
            fused = self.fuse_latents(full, video_ref, mask, vae, use_pca_channel_selection=use_pca_channel_selection,
                                      current_step=current_step, use_distill=use_distill,
                                      max_replace_threshold=max_replace_threshold)
            pred_x0 = fused[:, :, 1:]
        self.derivative_history.append(model_output)
        prev = (sample + dt * model_output).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return StepOutput(prev, pred_x0)

    def add_noise(self, original_samples, noise, timesteps, use_resample_sigma=False):
        sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
        schedule = self.timesteps.to(original_samples.device)
        idx = [self._index_for_timestep(t, schedule) for t in timesteps.to(original_samples.device)]
        sigma = sigmas[idx].flatten()
        while sigma.dim() < original_samples.dim():
            sigma = sigma.unsqueeze(-1)
        return (1.0 - sigma) * original_samples + sigma * noise


def timesteps_sigmas(sampling_steps: int, use_distill: bool = False, num_timesteps: int = 1000, num_distill: int = 50):
    """pipeline_longcat_video.py:316-331."""
    if use_distill:
        di = torch.arange(1, num_distill + 1, dtype=torch.float32)
        di = (di * (num_timesteps // num_distill)).round().long()
        ii = np.floor(np.linspace(0, num_distill, num=sampling_steps, endpoint=False)).astype(np.int64)
        sig = torch.flip(di, [0])[ii].float() / num_timesteps
        sig = sig - sig[-1]
    else:
        sig = torch.linspace(0.999, 0.000, sampling_steps)
    return sig.to(torch.float32)


def cfg_zero(noise_pred, guidance_scale):
    """[uncond, cond] batch -> CFG-zero combination, pipeline_longcat_video.py:875-885."""
    u, c = noise_pred.chunk(2)
    B = c.shape[0]
    pos, neg = c.reshape(B, -1), u.reshape(B, -1)
    st = (torch.sum(pos * neg, dim=1, keepdim=True) / (torch.sum(neg ** 2, dim=1, keepdim=True) + 1e-8)).view(B, 1, 1, 1, 1)
    return u * st + guidance_scale * (c - u * st)


def denoise_loop(dit, vae, scheduler, latents, prompt_embeds, prompt_attention_mask, num_inference_steps: int,
                 guidance_scale: float = 4.0, use_distill: bool = False, video_ref=None, mask=None, guided=False,
                 resample_steps=3, guide_steps=20, resample_round=20, omega=1.8, omega_resample=1.0,
                 use_pca_channel_selection=False, static=False, max_replace_threshold=None, generator=None,
                 do_cfg: bool = True, dit_dtype=torch.bfloat16, on_step=None):
    """generate_i2v's loop (:764-994).  latents [1,16,T,h,w] fp32 with the clean first frame in place; with CFG
    ``prompt_embeds`` is the [negative, positive] batch (:760-762).  Mutates and returns ``latents``."""
    device = latents.device
    scheduler.set_timesteps(num_inference_steps, sigmas=timesteps_sigmas(num_inference_steps, use_distill), device=device)
    timesteps = scheduler.timesteps
    if not hasattr(scheduler, "derivative_history"):
        scheduler.derivative_history = []
    for i, t in enumerate(timesteps):
        scheduler.derivative_history = []
        pred_x0, out = None, None
        for r in range(resample_steps if (guided and i < resample_round) else 1):
            if r > 0:
                scheduler.set_resample_mode(True)
                scheduler._step_index -= 1
            else:
                scheduler.set_resample_mode(False)
            t_dit = t.expand(latents.shape[0]).to(device=device, dtype=dit_dtype)
            x_in = (torch.cat([latents] * 2) if do_cfg else latents).to(dit_dtype)
            if do_cfg:
                t_dit = torch.cat([t_dit] * 2)
            ts = t_dit.unsqueeze(-1).repeat(1, x_in.shape[2])
            ts[:, :1] = 0
            v = dit(hidden_states=x_in, timestep=ts, encoder_hidden_states=prompt_embeds,
                    encoder_attention_mask=prompt_attention_mask, num_cond_latents=1)
            if do_cfg:
                v = cfg_zero(v, guidance_scale)
            v = -v
            out = scheduler.step(v[:, :, 1:], t, latents[:, :, 1:], video_ref=video_ref, mask=mask,
                                 guided=guided and i < guide_steps, resampling=r > 0, vae=vae,
                                 use_pca_channel_selection=use_pca_channel_selection, static=static, current_step=i,
                                 total_steps=len(timesteps), sample_full=latents, use_distill=use_distill,
                                 max_replace_threshold=max_replace_threshold, return_dict=True)
            if getattr(out, "pred_x0", None) is not None:
                pred_x0 = out.pred_x0
            if i >= resample_round:
                break
            if r < resample_steps - 1 and pred_x0 is not None:
                noise = torch.randn(pred_x0.shape, generator=generator).to(device=pred_x0.device, dtype=pred_x0.dtype)
                latents[:, :, 1:] = scheduler.add_noise(pred_x0, noise, t.expand(pred_x0.shape[0]).to(device=device),
                                                        use_resample_sigma=False)
        scheduler.set_resample_mode(False)
        if i < resample_round and len(scheduler.derivative_history) > 1 and guided:       # DSG (:946-986)
            w, g = scheduler.derivative_history[0], scheduler.derivative_history[-1]
            dims = list(range(1, g.dim()))
            dot = torch.sum(g * w, dim=dims, keepdim=True)
            ng = torch.sqrt(torch.sum(g ** 2, dim=dims, keepdim=True))
            nw = torch.sqrt(torch.sum(w ** 2, dim=dims, keepdim=True))
            cos = dot / (ng * nw + 1e-8)
            sin = torch.sin(torch.acos(torch.clamp(cos, -1.0, 1.0)))
            ratio = ng / (nw + 1e-8)
            om = omega_resample if i >= guide_steps else omega
            better = g + om * sin * (g - ratio * cos * w)
            scheduler._step_index -= 1
            b = scheduler.step(better, t, latents[:, :, 1:], guided=False, resampling=False, vae=vae, sample_full=latents,
                               use_distill=use_distill, return_dict=True)
            latents[:, :, 1:] = b.prev_sample
        elif out is not None:
            latents[:, :, 1:] = out.prev_sample
        if on_step is not None:
            on_step(i, latents)
    return latents


# ------------------------------------------------------------------------------------------------ refine (720p) pass
def refine_schedule(scheduler, num_inference_steps: int = 50, t_thresh: float = 0.5, device=None):
    """generate_refine step 4 (pipeline_longcat_video.py:1382-1391): the standard schedule cut at ``t_thresh``."""
    scheduler.set_timesteps(num_inference_steps, sigmas=timesteps_sigmas(num_inference_steps), device=device)
    timesteps = scheduler.timesteps
    if t_thresh:
        tt = torch.tensor(t_thresh * 1000, dtype=timesteps.dtype, device=timesteps.device)
        timesteps = torch.cat([tt.unsqueeze(0), timesteps[timesteps < tt]])
        scheduler.timesteps = timesteps
        scheduler.sigmas = torch.cat([timesteps / 1000, torch.zeros(1, device=timesteps.device)])
    return timesteps


def refine_padding(num_frames: int, num_cond_frames: int, temporal: int = 4, granularity: int = 4):
    """The BSA padding arithmetic of generate_refine (:1406-1421) -> (num_cond_latents, cond_frames_added,
    num_cond_frames_total, noise_frames_added)."""
    import math
    num_noise_frames = num_frames - num_cond_frames
    ncl = added = 0
    if num_cond_frames > 0:
        ncl = 1 + math.ceil((num_cond_frames - 1) / temporal)
        ncl = math.ceil(ncl / granularity) * granularity
        added = 1 + (ncl - 1) * temporal - num_cond_frames
        num_cond_frames = num_cond_frames + added
    nnl = math.ceil(num_noise_frames / temporal)
    nnl = math.ceil(nnl / granularity) * granularity
    return ncl, added, num_cond_frames, nnl * temporal - num_noise_frames


def _norm(vae, lat):
    z = vae.config.z_dim
    mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
    inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
    return (lat - mean) * inv_std


def refine_prepare(stage1_u8: torch.Tensor, image, vae, height: int, width: int, generator, t_thresh: float = 0.5,
                   num_cond_frames: int = 0, spatial_refine_only: bool = False, dtype=torch.bfloat16):
    """generate_refine step 5 (:1393-1455).  stage1_u8 uint8 [F, H0, W0, 3]; image: [1,3,height,width] in [-1,1] (already
    through video_processor.preprocess) or None.  -> (latents fp32 [1,16,T,h,w], num_cond_latents, cond_frames_added, new_frame_size)."""
    import torch.nn.functional as F
    nf = stage1_u8.shape[0]
    new_frames = nf if spatial_refine_only else 2 * nf
    s1 = stage1_u8.permute(0, 3, 1, 2).to(dtype=dtype)
    down = F.interpolate(s1, size=(height, width), mode="bilinear", align_corners=True)
    down = down.permute(1, 0, 2, 3).unsqueeze(0) / 255.0
    up = F.interpolate(down, size=(new_frames, height, width), mode="trilinear", align_corners=True)
    up = up * 2 - 1
    ncl, added, ncf, back = refine_padding(up.shape[2], num_cond_frames)
    up = torch.cat([up[:, :, 0:1].repeat(1, 1, added, 1, 1), up, up[:, :, -1:].repeat(1, 1, back, 1, 1)], dim=2)
    # the reference hands the bf16 clip to its bf16 VAE (run_longcat_worldforge_single.py:205); with a VAE of another dtype
    # the only meaningful reading is a cast to the VAE's dtype
    lat = _norm(vae, vae.encode(up.to(vae.dtype)).latent_dist.mode())
    noise = torch.randn(lat.shape, generator=generator, dtype=lat.dtype).to(lat.device)
    lat = (1 - t_thresh) * lat + t_thresh * noise
    latents = lat.to(torch.float32)                                           # prepare_latents(latents=latent_up, dtype=fp32)
    if image is not None:
        enc_in = image.to(dtype).unsqueeze(2)
        if added > 0:
            enc_in = torch.cat([enc_in[:, :, 0:1].repeat(1, 1, added, 1, 1), enc_in], dim=2)
        assert enc_in.shape[2] == ncf
        cond = _norm(vae, vae.encode(enc_in.to(vae.dtype)).latent_dist.mode().to(torch.float32))
        latents[:, :, : 1 + (ncf - 1) // 4] = cond
    return latents, ncl, added, new_frames


def refine_loop(dit, scheduler, latents, prompt_embeds, prompt_attention_mask, num_cond_latents: int, timesteps,
                dit_dtype=torch.bfloat16, on_step=None):
    """generate_refine's loop (:1467-1498): no CFG, no guidance; Euler steps on the noise latents only."""
    for i, t in enumerate(timesteps):
        x = latents.to(dit_dtype)
        ts = t.expand(x.shape[0]).to(dit_dtype).unsqueeze(-1).repeat(1, x.shape[2])
        ts[:, :num_cond_latents] = 0
        pred = dit(hidden_states=x, timestep=ts, encoder_hidden_states=prompt_embeds, encoder_attention_mask=prompt_attention_mask,
                   num_cond_latents=num_cond_latents)
        pred = -pred
        latents[:, :, num_cond_latents:] = scheduler.step(pred[:, :, num_cond_latents:], t, latents[:, :, num_cond_latents:],
                                                          return_dict=False)[0]
        if on_step is not None:
            on_step(i, latents)
    return latents


def vc_timesteps(scheduler, num_inference_steps: int, use_distill: bool = False, enhance_hf: bool = True, device=None):
    """generate_vc steps 4 (pipeline_longcat_video.py:1151-1164): the i2v schedule, with enhance_hf its tail below t = 500
    replaced by 10 uniform steps 500 -> 50 and the sigmas rebuilt as timesteps / 1000 (+ a final 0)."""
    import numpy as np
    assert not (use_distill and enhance_hf)
    scheduler.set_timesteps(num_inference_steps, sigmas=timesteps_sigmas(num_inference_steps, use_distill), device=device)
    timesteps = scheduler.timesteps
    if enhance_hf:
        tail = list(np.linspace(500, 0, 10, dtype=np.float32, endpoint=False))
        tail = [torch.tensor(t, device=device).unsqueeze(0) for t in tail]
        head = [t.unsqueeze(0) for t in timesteps if t > 500]
        timesteps = torch.cat(head + tail)
        scheduler.timesteps = timesteps
        scheduler.sigmas = torch.cat([timesteps / 1000, torch.zeros(1, device=timesteps.device)])
    return timesteps


def vc_loop(dit, scheduler, latents, prompt_embeds, prompt_attention_mask, num_cond_latents: int, timesteps,
            guidance_scale: float = 4.0, use_kv_cache: bool = True, do_cfg: bool = True, dit_dtype=torch.bfloat16, on_step=None):
    """generate_vc's loop (:1192-1250): the clean condition frames either go through the DiT once (no cross-attention) to
    fill the KV cache and only the noise frames are denoised against it, or ride along at timestep 0; CFG-zero, sign flip,
    plain Euler steps.  ``prompt_embeds`` is the [negative, positive] batch with CFG.  Returns the full latents."""
    kv = {}
    cond = None
    if use_kv_cache:
        cond = latents[:, :, :num_cond_latents]
        ts0 = torch.zeros(cond.shape[0], cond.shape[2]).to(device=latents.device, dtype=dit_dtype)
        empty = torch.zeros([cond.shape[0], 1, prompt_embeds.shape[2], prompt_embeds.shape[3]], device=latents.device, dtype=dit_dtype)
        _, kv = dit(hidden_states=cond.to(dit_dtype), timestep=ts0, encoder_hidden_states=empty, return_kv=True, skip_crs_attn=True)
        latents = latents[:, :, num_cond_latents:]
    for i, t in enumerate(timesteps):
        x = (torch.cat([latents] * 2) if do_cfg else latents).to(dit_dtype)
        ts = t.expand(x.shape[0]).to(dit_dtype).unsqueeze(-1).repeat(1, x.shape[2])
        if not use_kv_cache:
            ts[:, :num_cond_latents] = 0
        pred = dit(hidden_states=x, timestep=ts, encoder_hidden_states=prompt_embeds, encoder_attention_mask=prompt_attention_mask,
                   num_cond_latents=num_cond_latents, kv_cache_dict=kv)
        if do_cfg:
            pred = cfg_zero(pred, guidance_scale)
        pred = -pred
        if use_kv_cache:
            latents = scheduler.step(pred, t, latents, return_dict=False)[0]
        else:
            latents[:, :, num_cond_latents:] = scheduler.step(pred[:, :, num_cond_latents:], t, latents[:, :, num_cond_latents:],
                                                              return_dict=False)[0]
        if on_step is not None:
            on_step(i, latents)
    if use_kv_cache:
        latents = torch.cat([cond, latents], dim=2)
    return latents
