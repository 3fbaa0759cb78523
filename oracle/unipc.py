"""Oracle: the WorldForge UniPC flow scheduler with FLF fusion, restated on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``wan_for_worldforge/utils/scheduling_unipc_multistep_clean.py``:
set_timesteps :769-846 (flow sigmas :812-818), convert_model_output :925-976
(flow branch :952-958), multistep_uni_p_bh_update :978-1099, fuse_latents
:1248-1421, step :1423-1536, add_noise :1542-1585, resample tables :1594-1648.
Only the configuration WorldForge runs is restated: ``flow_prediction`` with
``use_flow_sigmas``, ``predict_x0``, ``bh2``, ``solver_order`` 2,
``lower_order_final``, ``final_sigmas_type='zero'``.  UniC
(:1101-1222) is never called by the reference pipeline and is not restated.

Every tensor expression keeps the reference's operand order and dtypes, because
from the first DSG step on the latents are bf16 (pipeline_wan_i2v_clean.py:708)
and each torch op then rounds to bf16: the CUDA kernels reproduce those
roundings and are compared bit-for-bit with this file.

The public mutable state the pipeline pokes (``_step_index``,
``lower_order_nums``, ``this_order``, ``model_outputs`` ...) keeps
the reference's names.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import flf as _flf


class StepOutput:
    def __init__(self, prev_sample, pred_x0):
        self.prev_sample = prev_sample
        self.pred_x0 = pred_x0

    def __getitem__(self, i):
        return (self.prev_sample, self.pred_x0)[i]


class OracleUniPC:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, flow_shift: float = 3.0):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                                      flow_shift=flow_shift, prediction_type="flow_prediction",
                                      use_flow_sigmas=True, predict_x0=True, solver_type="bh2",
                                      lower_order_final=True, final_sigmas_type="zero")
        self.num_inference_steps = None
        self.timesteps = None
        self.sigmas = None
        self.model_outputs = [None] * solver_order
        self.timestep_list = [None] * solver_order
        self.lower_order_nums = 0
        self.last_lower_order_nums = 0
        self.this_order = None
        self.last_this_order = None
        self.disable_corrector: List[int] = []
        self.last_sample = None
        self._step_index = None
        self._begin_index = None
        self.derivative_history = []
        self.resample_sigmas = None
        self.resample_timesteps = None
        self.is_resampling = False
        self.original_step_index = None
        self.flf_log = []          # (step, channels) each time FLF selection ran
        self.fuse_calls = 0

    step_index = property(lambda self: self._step_index)
    begin_index = property(lambda self: self._begin_index)

    # -- schedule (:769-846, :1594-1629) -----------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        shift = self.config.flow_shift
        alphas = np.linspace(1, 1 / n_train, num_inference_steps + 1)
        sig = 1.0 - alphas
        sig = np.flip(shift * sig / (1 + (shift - 1) * sig))[:-1].copy()
        timesteps = (sig * n_train).copy()
        sig = np.concatenate([sig, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self._begin_index = None
        # resample tables: the same sigmas minus the terminal zero; timestep = floor(sigma*1000)
        self.resample_sigmas = self.sigmas[:-1].clone()
        self.resample_timesteps = torch.floor(self.resample_sigmas * n_train).to(torch.int64)
        if device is not None:
            self.resample_sigmas = self.resample_sigmas.to(device)
            self.resample_timesteps = self.resample_timesteps.to(device)

    def set_resample_mode(self, enabled: bool):
        if enabled and not self.is_resampling:
            self.original_step_index = self._step_index
        self.is_resampling = enabled
        if not enabled and self.original_step_index is not None:
            self._step_index = self.original_step_index
            self.original_step_index = None

    def get_resample_timestep(self, step_index: int):
        if self.resample_timesteps is not None and step_index < len(self.resample_timesteps):
            return self.resample_timesteps[step_index].to(device=self.timesteps.device, dtype=self.timesteps.dtype)
        return self.timesteps[min(step_index, len(self.timesteps) - 1)]

    def _index_for_timestep(self, timestep, schedule=None):
        schedule = self.timesteps if schedule is None else schedule
        hits = (schedule == timestep).nonzero()
        if len(hits) == 0:
            return len(self.timesteps) - 1
        return hits[1 if len(hits) > 1 else 0].item()

    # -- sigma lookups: the resample tables mirror the ordinary ones ----------
    def _sigma_now(self):
        if self.is_resampling and self.resample_sigmas is not None:
            return self.resample_sigmas[min(self._step_index, len(self.resample_sigmas) - 1)]
        return self.sigmas[self._step_index]

    # -- x0 conversion (:937-958) -------------------------------------------
    def convert_model_output(self, model_output, sample=None):
        return sample - self._sigma_now() * model_output

    # -- UniP-bh2 predictor (:995-1099) ---------------------------------------
    def multistep_uni_p_bh_update(self, model_output, sample=None, order=None):
        m0 = self.model_outputs[-1]
        x = sample
        i = self._step_index
        if self.is_resampling and self.resample_sigmas is not None:
            n = len(self.resample_sigmas)
            sigma_t = self.sigmas[min(i + 1, n - 1)]
            sigma_s0 = self.resample_sigmas[min(i, n - 1)]
        else:
            sigma_t, sigma_s0 = self.sigmas[i + 1], self.sigmas[i]
        alpha_t, alpha_s0 = 1 - sigma_t, 1 - sigma_s0
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0

        D1 = None
        if order == 2:
            si = i - 1
            if self.is_resampling and self.resample_sigmas is not None:
                sigma_si = self.resample_sigmas[min(max(si, 0), len(self.resample_sigmas) - 1)]
            else:
                sigma_si = self.sigmas[si]
            lambda_si = torch.log(1 - sigma_si) - torch.log(sigma_si)
            rk = (lambda_si - lambda_s0) / h
            D1 = (self.model_outputs[-2] - m0) / rk
        elif order != 1:
            raise NotImplementedError("WorldForge runs solver_order 2")

        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = torch.expm1(hh)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if D1 is not None:
            rhos_p = torch.tensor([0.5], dtype=x.dtype, device=x.device)
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, torch.stack([D1], dim=1))
        else:
            pred_res = 0
        x_t = x_t_ - alpha_t * B_h * pred_res
        return x_t.to(x.dtype)

    # -- FLF fusion (:1248-1421) ----------------------------------------------
    def fuse_latents(self, pred_x0, video_ref, mask, vae=None, static=False, **kw):
        if mask is None or video_ref is None or vae is None:
            return pred_x0
        self.fuse_calls += 1
        dt, dev = pred_x0.dtype, pred_x0.device
        z = vae.config.z_dim
        mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(dev, dt)
        inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(dev, dt)
        lat = (pred_x0 / inv_std + mean).to(torch.float32)
        dec = vae.decode(lat, return_dict=False)[0]
        if video_ref.shape != dec.shape or mask.shape != (dec.shape[0], 1) + tuple(dec.shape[2:]):
            video_ref, mask = _flf.presize_guidance(video_ref, mask, dec.shape)       # :1300-1371
        ref = video_ref.to(dec.device, dec.dtype)
        m = mask.to(dec.device, dec.dtype)
        ref = 2.0 * ref - 1.0
        m = m.repeat(1, dec.shape[1], 1, 1, 1)
        fused = (ref * m + dec * (1 - m)).to(torch.float32)
        enc = vae.encode(fused).latent_dist.mode()
        enc = (enc - mean) * inv_std
        if kw.get("use_pca_channel_selection") and not kw.get("resampling", False):
            step = kw.get("current_step", 0)
            chans = _flf.select_channels(pred_x0, enc.to(dev, dt), step)
            self.flf_log.append((step, list(chans)))
            for c in chans:
                enc[:, c] = pred_x0[:, c]
        return enc.to(dev, dt)

    # -- one solver step (:1457-1536) -------------------------------------------
    def step(self, model_output, timestep, sample, return_dict=True, mask=None, guided=False,
             video_latents=None, resampling=False, vae=None, current_step=-1, resample_count=2,
             is_resample_round=False, static=False, **kw):
        if self._step_index is None:
            self._step_index = self._index_for_timestep(timestep) if self._begin_index is None else self._begin_index
        use_corrector = (self._step_index > 0 and self._step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        x0 = self.convert_model_output(model_output, sample=sample)
        if guided and video_latents is not None:
            x0 = self.fuse_latents(x0, video_latents, mask, vae=vae, current_step=current_step,
                                   total_steps=self.num_inference_steps, resampling=resampling,
                                   static=static, **kw)
        if not resampling:
            for j in range(self.config.solver_order - 1):
                self.model_outputs[j] = self.model_outputs[j + 1]
                self.timestep_list[j] = self.timestep_list[j + 1]
        self.model_outputs[-1] = x0
        self.timestep_list[-1] = timestep

        cap = min(self.config.solver_order, len(self.timesteps) - self._step_index)
        self.last_this_order = self.this_order
        self.this_order = min(cap, self.lower_order_nums + 1)
        assert self.this_order > 0
        if (not use_corrector) or (not is_resample_round) or resample_count < 2:
            self.last_sample = sample
        if resampling:
            self.derivative_history.append(model_output)
        prev = self.multistep_uni_p_bh_update(model_output=model_output, sample=sample, order=self.this_order)
        self.last_lower_order_nums = self.lower_order_nums
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return StepOutput(prev, x0)

    # -- IRR re-noise (:1542-1585) ------------------------------------------------
    def add_noise(self, original_samples, noise, timesteps, r=0, use_resample_sigma=False):
        if use_resample_sigma and self.resample_sigmas is not None:
            sigmas = self.resample_sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
            schedule = self.resample_timesteps.to(original_samples.device)
        else:
            sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
            schedule = self.timesteps.to(original_samples.device)
        timesteps = timesteps.to(original_samples.device)
        if self._begin_index is None:
            idx = [self._index_for_timestep(t, schedule) for t in timesteps]
        elif self._step_index is not None:
            idx = [min(self._step_index, len(sigmas) - 1) if use_resample_sigma else self._step_index] * timesteps.shape[0]
        else:
            idx = [self._begin_index] * timesteps.shape[0]
        sigma = sigmas[idx].flatten()
        while sigma.dim() < original_samples.dim():
            sigma = sigma.unsqueeze(-1)
        return (1 - sigma) * original_samples + sigma * noise
