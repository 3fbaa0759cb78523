"""LongCat-Video's guided i2v sampling (IRR + FLF + DSG) over the sm_100a engine.

Stands where ``FlowMatchEulerDiscreteScheduler`` (longcat_for_worldforge/longcat_video/modules/
scheduling_flow_match_euler_discrete.py:400-1233) and the loop of ``LongCatVideoPipeline.generate_i2v``
(longcat_video/pipeline_longcat_video.py:764-994) stand in the reference: same method names / keyword arguments /
public state (``_step_index``, ``derivative_history``, ``sigmas``, ``timesteps``), every tensor expression one fused kernel:

  pred_x0 = sample - sigma*v, prev = sample + dt*v     (:857, :900)         -> wf_x0_convert (twice, the second with -dt)
  CFG-zero + sign flip                                  (pipeline :875-888)  -> wf_cfg_zero
  re-noise (1-sigma)*x0 + sigma*n                       (:1068)              -> wf_renoise
  DSG                                                   (pipeline :946-971)  -> wf_dsg
  FLF: denorm, VAE decode, blend, VAE encode, normalise + channel replace (:1103-1222)
                                                        -> wf_latent_denorm, WfWanVAE, wf_flf_blend, wf_latent_norm_replace
  FLF scoring front end: per-channel min-max -> uint8   (:329-336, :147-151) -> wf_quantise_u8(mode 1); Farneback on the host

The latents and model outputs of this path are fp32 throughout (the DiT returns fp32, longcat_video_dit.py:365).  The
reference runs its VAE in bf16 (run_longcat_worldforge_single.py:205); the engine's VAE computes in tf32/fp32 and, to keep
the two boundary roundings the reference has, can round the decoder input and the blended clip to bf16
(``vae_io_bf16=True``; off by default, i.e. an fp32 VAE as in the Wan path).
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import lib
from .flf_select import _flow_pair
from .scheduler import latent_stats


def flow_similarity(ref: torch.Tensor, cand: torch.Tensor) -> float:
    """LongCat's metric mix (:205-244): outliers by OR, weights 0.4 / 0.4 / 0.2.  Inputs [T-1, 2, H, W]."""
    d = ref - cand
    epe = torch.sqrt((d ** 2).sum(dim=1) + 1e-8)
    dot = (ref * cand).sum(dim=1)
    rn = torch.sqrt((ref ** 2).sum(dim=1) + 1e-8)
    cn = torch.sqrt((cand ** 2).sum(dim=1) + 1e-8)
    ang = torch.acos(torch.clamp(dot / (rn * cn + 1e-8), -1.0, 1.0)) * 180.0 / torch.pi
    outlier = (epe > 3.0) | (epe > rn * 0.05)
    err = (0.4 * torch.clamp(epe.mean() / 10.0, 0.0, 1.0) + 0.4 * torch.clamp(outlier.float().mean() / 0.5, 0.0, 1.0)
           + 0.2 * torch.clamp(ang.mean() / 30.0, 0.0, 1.0))
    return torch.clamp(1.0 - err, 0.0, 1.0).item()


def selection_policy(scores, step: int, use_distill: bool, max_replace_threshold: Optional[int]) -> List[int]:
    """(:338-379) early steps: the single lowest score; later: below mean - 0.625 std, at least 1, at most the cap."""
    s = np.array(scores)
    if step <= (3 if use_distill else 5):
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


def similarity_from_means(epe: float, outlier: float, ang: float) -> float:
    """LongCat's score from the three per-channel means (:229-244): weights 0.4 / 0.4 / 0.2."""
    clamp = lambda v: min(max(v, 0.0), 1.0)
    return clamp(1.0 - (0.4 * clamp(epe / 10.0) + 0.4 * clamp(outlier / 0.5) + 0.2 * clamp(ang / 30.0)))


class LongCatChannelSelector:
    def __init__(self, threads: int = 0, device_flow=None):
        self.threads = threads or min(32, os.cpu_count() or 8)
        self._pool = None
        self.last_scores = None
        # Farneback + the flow metrics on the GPU (wf_farneback_u8 / wf_flow_metrics with the OR-ed outlier test) for frame
        # sizes the device path covers; WF_FLF_GPU=0 keeps OpenCV on host threads
        self.device_flow = (os.environ.get("WF_FLF_GPU", "1") == "1") if device_flow is None else bool(device_flow)

    def _flows(self, u8: np.ndarray) -> torch.Tensor:
        C, T = u8.shape[:2]
        jobs = [(u8[c, t], u8[c, t + 1]) for c in range(C) for t in range(T - 1)]
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.threads)
        res = list(self._pool.map(_flow_pair, jobs, chunksize=4))
        fl = np.stack(res, axis=0).reshape(C, T - 1, *res[0].shape).transpose(0, 1, 4, 2, 3)
        return torch.from_numpy(np.ascontiguousarray(fl)).float()

    def select(self, pred_x0, enc, step: int, use_distill: bool, max_replace_threshold) -> List[int]:
        if step < 2:
            return []
        C = pred_x0.shape[1]
        q = torch.empty((2,) + tuple(pred_x0.shape[1:]), dtype=torch.uint8, device=pred_x0.device)
        for c in range(C):                    # per-channel min-max (:329-336)
            lib.quantise_u8(enc[0, c], mode=1, out=q[0, c])
            lib.quantise_u8(pred_x0[0, c], mode=1, out=q[1, c])
        from .flf_select import FlowChannelSelector
        if self.device_flow and FlowChannelSelector.device_path_covers(*q.shape[-2:]):
            flows = lib.farneback_u8(q.reshape((2 * C,) + tuple(q.shape[2:])))          # [2C, T-1, H, W, 2] on the device
            means = lib.flow_metrics(flows[:C].contiguous(), flows[C:].contiguous(), outlier_or=True).cpu()
            self.last_scores = [similarity_from_means(float(means[c, 0]), float(means[c, 1]), float(means[c, 2])) for c in range(C)]
            return selection_policy(self.last_scores, step, use_distill, max_replace_threshold)
        both = q.cpu().numpy()
        ref_fl, pred_fl = self._flows(both[0]), self._flows(both[1])
        self.last_scores = [flow_similarity(ref_fl[c], pred_fl[c]) for c in range(C)]
        return selection_policy(self.last_scores, step, use_distill, max_replace_threshold)


class StepOutput:
    def __init__(self, prev_sample, pred_x0=None):
        self.prev_sample, self.pred_x0 = prev_sample, pred_x0


class WfFlowMatchEulerScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, vae_io_bf16: bool = False, **unused):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift, stochastic_sampling=False)
        self.shift = shift
        self.vae_io_bf16 = vae_io_bf16
        self._step_index = None
        self._begin_index = None
        self.derivative_history = []
        self.is_resampling = False
        self.flf_log = []
        self.fuse_calls = 0
        self._selector = None

    step_index = property(lambda self: self._step_index)
    begin_index = property(lambda self: self._begin_index)

    # generate_refine assigns scheduler.timesteps / scheduler.sigmas directly (pipeline_longcat_video.py:1390-1391): the
    # host copies the step arithmetic reads follow such assignments
    @property
    def timesteps(self):
        return self._timesteps

    @timesteps.setter
    def timesteps(self, value):
        self._timesteps = value
        self._timesteps_host = value.detach().to("cpu", torch.float32)

    @property
    def sigmas(self):
        return self._sigmas

    @sigmas.setter
    def sigmas(self, value):
        self._sigmas = value
        self._sigmas_host = value.detach().to("cpu", torch.float32)

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, **kw):
        if sigmas is None:
            raise NotImplementedError("WorldForge's LongCat pipeline always passes its own sigmas (pipeline_longcat_video.py:765-766)")
        sig = (sigmas.detach().cpu().numpy() if isinstance(sigmas, torch.Tensor) else np.asarray(sigmas)).astype(np.float32)
        sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        host = torch.from_numpy(sig).to(torch.float32)
        sig_h, ts_h = torch.cat([host, torch.zeros(1)]), host * self.config.num_train_timesteps
        self.sigmas = sig_h.to(device) if device is not None else sig_h
        self.timesteps = ts_h.to(device) if device is not None else ts_h
        self.num_inference_steps = len(sig)
        self._step_index = None
        self._begin_index = None
        self.derivative_history = []
        self._selector = None

    def set_resample_mode(self, enabled: bool):
        self.is_resampling = enabled

    def _index_for_timestep(self, timestep) -> int:
        t = float(timestep)
        idx = (self._timesteps_host == t).nonzero()
        return idx[1 if len(idx) > 1 else 0].item()

    # ------------------------------------------------------------------------------------------ FLF (:1072-1233)
    def fuse_latents(self, pred_original_sample, video_latents, mask, vae, use_pca_channel_selection=False, static=False,
                     current_step=0, total_steps=50, use_distill=False, max_replace_threshold=None):
        if mask is None or video_latents is None or vae is None:
            return pred_original_sample
        self.fuse_calls += 1
        x0 = pred_original_sample.contiguous()
        mean_h, inv_std_h = latent_stats(vae.config.latents_mean, vae.config.latents_std, x0.dtype)
        z = lib.latent_denorm(x0, mean_h, inv_std_h)
        if self.vae_io_bf16:
            z = z.to(torch.bfloat16).to(torch.float32)
        dec = vae.decode(z, return_dict=False)[0]
        if tuple(video_latents.shape) != tuple(dec.shape):
            raise ValueError(f"Dimension mismatch! decoded_video={tuple(dec.shape)}, video_ref={tuple(video_latents.shape)}")
        fused = lib.flf_blend(dec.contiguous(), video_latents.to(torch.float32).contiguous(), mask.to(torch.float32).contiguous())
        if self.vae_io_bf16:
            fused = fused.to(torch.bfloat16).to(torch.float32)
        enc = vae.encode(fused).latent_dist.mode().contiguous()
        chans: List[int] = []
        if use_pca_channel_selection:
            if current_step >= 2:
                enc_n = lib.latent_norm_replace(enc, x0, mean_h, inv_std_h, [])
                if self._selector is None:
                    self._selector = LongCatChannelSelector()
                chans = self._selector.select(x0, enc_n, current_step, use_distill, max_replace_threshold)
            self.flf_log.append((current_step, list(chans)))
        return lib.latent_norm_replace(enc, x0, mean_h, inv_std_h, chans)

    # ----------------------------------------------------------------------------------------- step (:740-912)
    def step(self, model_output, timestep, sample, return_dict=True, video_ref=None, mask=None, guided=False,
             resampling=False, vae=None, use_pca_channel_selection=False, static=False, current_step=-1, total_steps=50,
             sample_full=None, use_distill=False, max_replace_threshold=None, **kw):
        if self._step_index is None:
            self._step_index = self._index_for_timestep(timestep) if self._begin_index is None else self._begin_index
        sample = sample.to(torch.float32).contiguous()
        v = model_output.contiguous()
        sigma = float(self._sigmas_host[self._step_index])
        dt = float(self._sigmas_host[self._step_index + 1] - self._sigmas_host[self._step_index])
        pred_x0 = lib.x0_convert(sample, v, sigma)
        if guided and video_ref is not None and not resampling and sample_full is not None:
            v_full = torch.cat([torch.zeros_like(v[:, :, 0:1]), v], dim=2)
            full = lib.x0_convert(sample_full.to(torch.float32).contiguous(), v_full, sigma)
            fused = self.fuse_latents(full, video_ref, mask, vae, use_pca_channel_selection=use_pca_channel_selection,
                                      static=static, current_step=current_step, total_steps=total_steps,
                                      use_distill=use_distill, max_replace_threshold=max_replace_threshold)
            pred_x0 = fused[:, :, 1:]
        self.derivative_history.append(model_output)
        prev = lib.x0_convert(sample, v, -dt)              # sample + dt*v  ==  sample - (-dt)*v, same two roundings
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return StepOutput(prev, pred_x0)

    def add_noise(self, original_samples, noise, timesteps, use_resample_sigma: bool = False):
        idx = self._index_for_timestep(timesteps.reshape(-1)[0])
        sig = self._sigmas_host[idx]
        x0 = original_samples.to(torch.float32).contiguous()
        return lib.renoise(x0, noise.to(torch.float32).contiguous(), float(1.0 - sig), float(sig))


def timesteps_sigmas(sampling_steps: int, use_distill: bool = False, num_timesteps: int = 1000, num_distill: int = 50):
    """LongCatVideoPipeline.get_timesteps_sigmas (pipeline_longcat_video.py:316-331)."""
    if use_distill:
        di = torch.arange(1, num_distill + 1, dtype=torch.float32)
        di = (di * (num_timesteps // num_distill)).round().long()
        ii = np.floor(np.linspace(0, num_distill, num=sampling_steps, endpoint=False)).astype(np.int64)
        sig = torch.flip(di, [0])[ii].float() / num_timesteps
        sig = sig - sig[-1]
    else:
        sig = torch.linspace(0.999, 0.000, sampling_steps)
    return sig.to(torch.float32)


@torch.no_grad()
def denoise_loop(dit, vae, scheduler, latents, prompt_embeds, prompt_attention_mask, num_inference_steps: int,
                 guidance_scale: float = 4.0, use_distill: bool = False, video_ref=None, mask=None, guided: bool = False,
                 resample_steps: int = 3, guide_steps: int = 20, resample_round: int = 20, omega: float = 1.8,
                 omega_resample: float = 1.0, use_pca_channel_selection: bool = False, static: bool = False,
                 max_replace_threshold=None, generator=None, do_cfg: bool = True, on_step=None):
    """The loop of generate_i2v (:764-994).  latents [1,16,T,h,w] fp32 on the device, clean first frame in place; with CFG
    ``prompt_embeds`` is the [negative, positive] batch (:760-762).  Mutates and returns ``latents``."""
    device = latents.device
    dit_dtype = dit.dtype
    scheduler.set_timesteps(num_inference_steps, sigmas=timesteps_sigmas(num_inference_steps, use_distill), device=device)
    timesteps = scheduler.timesteps
    if video_ref is not None and guided:
        video_ref = video_ref.to(device=device, dtype=torch.float32)
    if mask is not None and guided:
        mask = mask.to(device)
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
                v = lib.cfg_zero(v[1:2].contiguous(), v[0:1].contiguous(), guidance_scale)      # includes the sign flip
            else:
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
                noise = torch.randn(pred_x0.shape, generator=generator).pin_memory().to(device=device, non_blocking=True)
                latents[:, :, 1:] = scheduler.add_noise(pred_x0, noise, t.expand(pred_x0.shape[0]).to(device=device))
        scheduler.set_resample_mode(False)
        if i < resample_round and len(scheduler.derivative_history) > 1 and guided:
            w, g = scheduler.derivative_history[0], scheduler.derivative_history[-1]
            better = lib.dsg(g.contiguous(), w.contiguous(), omega_resample if i >= guide_steps else omega)
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
    """BSA padding of generate_refine (:1406-1421) -> (num_cond_latents, cond_frames_added, num_cond_frames, noise_frames_added)."""
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


@torch.no_grad()
def refine_prepare(stage1_u8: torch.Tensor, image, vae, height: int, width: int, generator, t_thresh: float = 0.5,
                   num_cond_frames: int = 0, spatial_refine_only: bool = False, device="cuda"):
    """generate_refine step 5 (:1393-1455).  stage1_u8: uint8 [F,H0,W0,3] (the stage-1 frames); image: [1,3,height,width] in
    [-1,1] (after video_processor.preprocess) or None.  -> (latents fp32 [1,16,T,h,w] on the device, num_cond_latents,
    cond_frames_added, new_frame_size).  The resize chain is one kernel (wf_refine_upsample); the noise comes from the CPU
    generator as in the reference (:1427)."""
    nf = stage1_u8.shape[0]
    new_frames = nf if spatial_refine_only else 2 * nf
    ncl, added, ncf, back = refine_padding(new_frames, num_cond_frames)
    up = lib.refine_upsample(stage1_u8.to(device).contiguous(), new_frames, height, width, added, back).unsqueeze(0)
    mean_h, inv_std_h = latent_stats(vae.config.latents_mean, vae.config.latents_std, torch.float32)
    enc = vae.encode(up).latent_dist.mode().contiguous()
    lat = lib.latent_norm_replace(enc, enc, mean_h, inv_std_h, [])
    noise = torch.randn(lat.shape, generator=generator, dtype=lat.dtype).pin_memory().to(device=device, non_blocking=True)
    latents = lib.renoise(lat, noise, float(1 - t_thresh), float(t_thresh))
    if image is not None:
        enc_in = image.to(device=device, dtype=torch.float32).unsqueeze(2)
        if added > 0:
            enc_in = torch.cat([enc_in[:, :, 0:1].repeat(1, 1, added, 1, 1), enc_in], dim=2)
        assert enc_in.shape[2] == ncf
        cenc = vae.encode(enc_in.contiguous()).latent_dist.mode().contiguous()
        latents[:, :, : 1 + (ncf - 1) // 4] = lib.latent_norm_replace(cenc, cenc, mean_h, inv_std_h, [])
    return latents, ncl, added, new_frames


@torch.no_grad()
def refine_loop(dit, scheduler, latents, prompt_embeds, prompt_attention_mask, num_cond_latents: int, timesteps, on_step=None):
    """generate_refine's loop (:1467-1498): one DiT forward (block-sparse self-attention when dit.enable_bsa() was called),
    sign flip, Euler step on the noise latents only; no CFG, no IRR / FLF / DSG."""
    dit_dtype = dit.dtype
    for i, t in enumerate(timesteps):
        x = latents.to(dit_dtype)
        ts = t.expand(x.shape[0]).to(dit_dtype).unsqueeze(-1).repeat(1, x.shape[2])
        ts[:, :num_cond_latents] = 0
        pred = dit(hidden_states=x, timestep=ts, encoder_hidden_states=prompt_embeds, encoder_attention_mask=prompt_attention_mask,
                   num_cond_latents=num_cond_latents)
        # -pred, then sample + dt*(-pred): x0_convert(sample, pred, dt) = sample - dt*pred has the same two roundings
        noise_pred = -pred
        latents[:, :, num_cond_latents:] = scheduler.step(noise_pred[:, :, num_cond_latents:], t, latents[:, :, num_cond_latents:],
                                                          return_dict=False)[0]
        if on_step is not None:
            on_step(i, latents)
    return latents


# ------------------------------------------------------------------------------------ video continuation (KV cache)
def vc_timesteps(scheduler, num_inference_steps: int, use_distill: bool = False, enhance_hf: bool = True, device=None):
    """generate_vc step 4 (pipeline_longcat_video.py:1151-1164): the i2v schedule; with ``enhance_hf`` its part below t = 500
    is replaced by 10 uniform steps 500 -> 50 and the sigmas are rebuilt as timesteps / 1000 plus a final 0."""
    if use_distill and enhance_hf:
        raise ValueError("use_distill and enhance_hf cannot both be True")
    scheduler.set_timesteps(num_inference_steps, sigmas=timesteps_sigmas(num_inference_steps, use_distill), device=device)
    timesteps = scheduler.timesteps
    if enhance_hf:
        tail = [torch.tensor(t, device=device).unsqueeze(0) for t in np.linspace(500, 0, 10, dtype=np.float32, endpoint=False)]
        head = [t.unsqueeze(0) for t in timesteps if t > 500]
        timesteps = torch.cat(head + tail)
        scheduler.timesteps = timesteps
        scheduler.sigmas = torch.cat([timesteps / 1000, torch.zeros(1, device=timesteps.device)])
    return timesteps


def vc_loop(dit, scheduler, latents, prompt_embeds, prompt_attention_mask, num_cond_latents: int, timesteps,
            guidance_scale: float = 4.0, use_kv_cache: bool = True, offload_kv_cache: bool = False, do_cfg: bool = True, on_step=None):
    """The denoising loop of generate_vc (pipeline_longcat_video.py:1192-1250, LongCat's long-video continuation): the
    clean latents of the condition frames go through the DiT once, without cross-attention, to fill every layer's K/V cache
    (_cache_clean_latents, :336-350); each step then runs the noise frames alone against [cache | own] keys.  Without the
    cache the condition frames ride along at timestep 0.  CFG-zero + sign flip (``wf_cfg_zero``), plain Euler steps.
    latents [1,16,T,h,w] fp32 on the device with the clean condition latents in front; returns the full latents."""
    device, dit_dtype = latents.device, dit.dtype
    kv, cond = {}, None
    if use_kv_cache:
        cond = latents[:, :, :num_cond_latents]
        ts0 = torch.zeros(cond.shape[0], cond.shape[2], device=device, dtype=dit_dtype)
        empty = torch.zeros([cond.shape[0], 1, prompt_embeds.shape[2], prompt_embeds.shape[3]], device=device, dtype=dit_dtype)
        _, kv = dit(hidden_states=cond.to(dit_dtype), timestep=ts0, encoder_hidden_states=empty, return_kv=True, skip_crs_attn=True,
                    offload_kv_cache=offload_kv_cache)
        latents = latents[:, :, num_cond_latents:].contiguous()
    for i, t in enumerate(timesteps):
        x = (torch.cat([latents] * 2) if do_cfg else latents).to(dit_dtype)
        ts = t.expand(x.shape[0]).to(device=device, dtype=dit_dtype).unsqueeze(-1).repeat(1, x.shape[2])
        if not use_kv_cache:
            ts[:, :num_cond_latents] = 0
        pred = dit(hidden_states=x, timestep=ts, encoder_hidden_states=prompt_embeds, encoder_attention_mask=prompt_attention_mask,
                   num_cond_latents=num_cond_latents, kv_cache_dict=kv)
        pred = lib.cfg_zero(pred[1:2].contiguous(), pred[0:1].contiguous(), guidance_scale) if do_cfg else -pred   # sign flip included
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
