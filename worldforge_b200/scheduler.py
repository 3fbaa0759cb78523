"""WorldForge's UniPC flow scheduler with FLF fusion, driving the fused sm_100a kernels.

Drop-in for the object the reference pipeline installs as ``pipe.scheduler``
(reference wan_for_worldforge/infer_worldforge.py:201-202; class
``UniPCMultistepScheduler`` in utils/scheduling_unipc_multistep_clean.py:649-1648): same
method names, same keyword arguments, same PUBLIC MUTABLE STATE that the pipeline pokes
between calls (``_step_index``, ``lower_order_nums``, ``last_lower_order_nums``,
``this_order``, ``last_this_order``, ``last_sample``, ``model_outputs``,
``derivative_history``, ``disable_corrector``; pipeline_wan_i2v_clean.py:556-706).

Split of work: the scalar logic of the solver (sigma tables, orders, the handful of fp32
coefficients per step) stays on the host, evaluated with the same fp32 operation order as the
reference so the coefficients are bit-identical; every tensor expression is ONE fused kernel
launch through the C ABI (worldforge_b200.lib):

  convert_model_output        -> wf_x0_convert           (:952-958)
  multistep_uni_p_bh_update   -> wf_unip_update          (:1084-1099)
  add_noise                   -> wf_renoise              (:1584)
  fuse_latents                -> wf_latent_denorm, VAE decode, wf_flf_blend, VAE encode,
                                 wf_quantise_u8 + host Farneback, wf_latent_norm_replace (:1248-1421)

Only the configuration WorldForge uses is implemented (flow_prediction, flow sigmas,
predict_x0, bh2, solver_order 2, lower_order_final, final sigma zero); UniC is dead code in the
reference pipeline and is not provided.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import flf_select, lib


class SchedulerStepOutput:
    """``prev_sample`` / ``pred_x0`` pair (the reference's CustomSchedulerOutput, :1522-1536)."""

    def __init__(self, prev_sample, pred_x0):
        self.prev_sample, self.pred_x0 = prev_sample, pred_x0

    def __getitem__(self, i):
        return (self.prev_sample, self.pred_x0)[i]


def latent_stats(latents_mean, latents_std, dtype):
    """Per-channel mean and 1/std as host floats ROUNDED to ``dtype`` the way the reference builds
    them (``torch.tensor(mean).to(dtype)``, ``1.0 / torch.tensor(std).to(dtype)``; :1272-1279)."""
    mean = torch.tensor(latents_mean).to(dtype)
    inv_std = 1.0 / torch.tensor(latents_std).to(dtype)
    return mean.float().tolist(), inv_std.float().tolist()


def _round_to(v: float, dtype) -> float:
    """v rounded to ``dtype`` (bf16) and widened back - what a device-side 0-dim fp32 tensor becomes when torch
    casts it to the common dtype of an op with a bf16 tensor."""
    if dtype == torch.bfloat16:
        return float(torch.tensor(v, dtype=torch.float32).to(torch.bfloat16))
    return float(v)


def unip_coefficients(sigmas: torch.Tensor, resample_sigmas: Optional[torch.Tensor], step_index: int, order: int,
                      resampling: bool):
    """(c_x, c_m0, r_1, c_res) of the UniP-bh2 predictor as fp32 values (:1005-1061,1084-1089).

    ``x' = c_x*x - c_m0*m0 - c_res*0.5*(m1-m0)/r_1`` with c_x = sigma_t/sigma_s0,
    c_m0 = alpha_t*expm1(-h), c_res = alpha_t*B_h (B_h = expm1(-h) for bh2).  fp32 0-dim torch
    tensors are used on purpose: same operations, same order, same rounding as the reference.
    """
    i = step_index
    if resampling and resample_sigmas is not None:
        n = len(resample_sigmas)
        sigma_t = sigmas[min(i + 1, n - 1)]
        sigma_s0 = resample_sigmas[min(i, n - 1)]
    else:
        sigma_t, sigma_s0 = sigmas[i + 1], sigmas[i]
    alpha_t, alpha_s0 = 1 - sigma_t, 1 - sigma_s0
    lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
    lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
    h = lambda_t - lambda_s0
    rk = torch.tensor(1.0)
    if order == 2:
        si = i - 1
        if resampling and resample_sigmas is not None:
            sigma_si = resample_sigmas[min(max(si, 0), len(resample_sigmas) - 1)]
        else:
            sigma_si = sigmas[si]
        lambda_si = torch.log(1 - sigma_si) - torch.log(sigma_si)
        rk = (lambda_si - lambda_s0) / h
    hh = -h
    h_phi_1 = torch.expm1(hh)
    B_h = torch.expm1(hh)
    return (float(sigma_t / sigma_s0), float(alpha_t * h_phi_1), float(rk), float(alpha_t * B_h))


def unip_kernel_args(co, order: int, resampling: bool, x_dtype, m0_dtype, m1_dtype):
    """Turn the fp32 coefficients into the arguments of wf_unip_update, reproducing how torch's CUDA
    kernels treat the scalar operand of each tensor-scalar op of the reference:

    * not resampling: every coefficient is built from ``self.sigmas`` (host tensors, :838), i.e. a HOST
      scalar - kept in fp32 inside the kernel; ``tensor / host_scalar`` is evaluated as
      ``tensor * (1/host_scalar)``;
    * resampling: sigma_s0 comes from ``self.resample_sigmas``, which the reference moves to the device
      (:843-846), so every coefficient is a DEVICE 0-dim tensor - cast to the op's common dtype (rounded
      to bf16 when the other operand is bf16) and ``tensor / device_scalar`` is a true division.
    """
    c_x, c_m0, rk, c_res = co
    bf = torch.bfloat16
    if not resampling:
        return (c_x, c_m0, float(torch.tensor(1.0) / torch.tensor(rk, dtype=torch.float32)) if order == 2 else 1.0, True, c_res)
    diff_dt = bf if (order == 2 and m0_dtype == bf and m1_dtype == bf) else torch.float32
    pred_dt = bf if (diff_dt == bf and x_dtype == bf) else torch.float32
    return (_round_to(c_x, x_dtype), _round_to(c_m0, m0_dtype), _round_to(rk, diff_dt), False, _round_to(c_res, pred_dt))


class WfUniPCScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, flow_shift: float = 3.0, **unused):
        if solver_order != 2:
            raise NotImplementedError("WorldForge runs UniPC with solver_order=2")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                                      flow_shift=flow_shift, prediction_type="flow_prediction", use_flow_sigmas=True,
                                      predict_x0=True, solver_type="bh2", lower_order_final=True,
                                      final_sigmas_type="zero")
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
        self.resample_sigmas = None       # host copy (fp32); the reference keeps it on the device
        self.resample_timesteps = None
        self.is_resampling = False
        self.original_step_index = None
        self.flf_log = []                 # (step, channels) per FLF selection, for parity tests
        self.fuse_calls = 0
        self._selector = None
        self._presized = None             # (key, clip, mask, (resized clip, resized mask)) of the last mis-sized guidance pair

    @classmethod
    def from_config(cls, config, **kw):
        cfg = dict(config) if not isinstance(config, SimpleNamespace) else vars(config)
        cfg.update(kw)
        return cls(num_train_timesteps=cfg.get("num_train_timesteps", 1000), solver_order=cfg.get("solver_order", 2),
                   flow_shift=cfg.get("flow_shift", 3.0))

    step_index = property(lambda self: self._step_index)
    begin_index = property(lambda self: self._begin_index)

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def scale_model_input(self, sample, *a, **k):
        return sample

    def __len__(self):
        return self.config.num_train_timesteps

    # ----------------------------------------------------------------- schedule (:769-846, :1594-1629)
    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train, shift = self.config.num_train_timesteps, self.config.flow_shift
        alphas = np.linspace(1, 1 / n_train, num_inference_steps + 1)
        sig = 1.0 - alphas
        sig = np.flip(shift * sig / (1 + (shift - 1) * sig))[:-1].copy()
        timesteps = (sig * n_train).copy()
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self._timesteps_host = self.timesteps.cpu()
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self._begin_index = None
        self._selector = None
        self._presized = None
        self.resample_sigmas = self.sigmas[:-1].clone()
        self.resample_timesteps = torch.floor(self.resample_sigmas * n_train).to(torch.int64)
        self._resample_timesteps_dev = self.resample_timesteps.to(device) if device is not None else self.resample_timesteps

    def set_resample_mode(self, enabled: bool):
        if enabled and not self.is_resampling:
            self.original_step_index = self._step_index
        self.is_resampling = enabled
        if not enabled and self.original_step_index is not None:
            self._step_index = self.original_step_index
            self.original_step_index = None

    def get_resample_timestep(self, step_index: int) -> torch.Tensor:
        if self.resample_timesteps is not None and step_index < len(self.resample_timesteps):
            return self._resample_timesteps_dev[step_index].to(dtype=self.timesteps.dtype)
        return self.timesteps[min(step_index, len(self.timesteps) - 1)]

    def _index_for_timestep(self, timestep, schedule_host=None) -> int:
        schedule = self._timesteps_host if schedule_host is None else schedule_host
        t = int(timestep)
        hits = (schedule == t).nonzero()
        if len(hits) == 0:
            return len(self.timesteps) - 1
        return hits[1 if len(hits) > 1 else 0].item()

    def _sigma_now(self) -> torch.Tensor:
        if self.is_resampling and self.resample_sigmas is not None:
            return self.resample_sigmas[min(self._step_index, len(self.resample_sigmas) - 1)]
        return self.sigmas[self._step_index]

    # ------------------------------------------------------------------------------ tensor ops -> kernels
    def convert_model_output(self, model_output, *args, sample=None, **kw):
        sigma = float(self._sigma_now())
        if self.is_resampling and self.resample_sigmas is not None:
            sigma = _round_to(sigma, model_output.dtype)     # device-side scalar: cast to the tensor's dtype
        return lib.x0_convert(sample.contiguous(), model_output.contiguous(), sigma)

    def multistep_uni_p_bh_update(self, model_output=None, *args, sample=None, order=None, **kw):
        resampling = self.is_resampling and self.resample_sigmas is not None
        co = unip_coefficients(self.sigmas, self.resample_sigmas, self._step_index, order, resampling)
        m0 = self.model_outputs[-1]
        m1 = self.model_outputs[-2] if order == 2 else None
        args = unip_kernel_args(co, order, resampling, sample.dtype, m0.dtype, m1.dtype if m1 is not None else None)
        return lib.unip_update(sample.contiguous(), m0, m1, order, *args)

    def add_noise(self, original_samples, noise, timesteps, r: int = 0, use_resample_sigma: bool = False):
        if use_resample_sigma and self.resample_sigmas is not None:
            sigmas, schedule = self.resample_sigmas, self.resample_timesteps
        else:
            sigmas, schedule = self.sigmas, self._timesteps_host
        nt = int(timesteps.numel())
        if nt != 1:
            raise NotImplementedError("batch size 1 (as in the entry script)")
        if self._begin_index is None:
            idx = self._index_for_timestep(timesteps.reshape(-1)[0], schedule)
        elif self._step_index is not None:
            idx = min(self._step_index, len(sigmas) - 1) if use_resample_sigma else self._step_index
        else:
            idx = self._begin_index
        dt = original_samples.dtype
        sig = sigmas[idx].to(dt).view(1)             # sigma held as a 1-element tensor of x0's dtype
        return lib.renoise(original_samples.contiguous(), noise.contiguous(), float((1 - sig).float()),
                           float(sig.float()))

    # ----------------------------------------------------------------------------------- FLF (:1248-1421)
    def fuse_latents(self, pred_original_sample, video_latents, mask, vae=None, static=False, **kw):
        if mask is None or video_latents is None or vae is None:
            return pred_original_sample
        self.fuse_calls += 1
        x0 = pred_original_sample.contiguous()
        mean_h, inv_std_h = latent_stats(vae.config.latents_mean, vae.config.latents_std, x0.dtype)
        z = lib.latent_denorm(x0, mean_h, inv_std_h)
        with lib.phase("flf.vae_decode"):
            dec = vae.decode(z, return_dict=False)[0]
        if tuple(video_latents.shape) != tuple(dec.shape) or tuple(mask.shape) != (dec.shape[0], 1) + tuple(dec.shape[2:]):
            # the reference resizes a mis-sized clip / mask on every call (:1300-1371); here once per (clip, mask) pair
            key = (video_latents.data_ptr(), video_latents._version, mask.data_ptr(), mask._version, tuple(dec.shape))
            if self._presized is None or self._presized[0] != key:
                from . import inputs
                sized = inputs.presize_guidance(video_latents.to(dec.device), mask.to(dec.device), dec.shape)
                self._presized = (key, video_latents, mask, sized)       # the sources are held: their addresses stay theirs
            video_latents, mask = self._presized[3]
        ref = video_latents if video_latents.dtype == torch.float32 else video_latents.to(torch.float32)
        m = mask if mask.dtype == torch.float32 else mask.to(torch.float32)
        fused = lib.flf_blend(dec.contiguous(), ref.contiguous(), m.contiguous())
        with lib.phase("flf.vae_encode"):
            enc = vae.encode(fused).latent_dist.mode().contiguous()
        chans: List[int] = []
        if kw.get("use_pca_channel_selection") and not kw.get("resampling", False):
            step = kw.get("current_step", 0)
            if step >= 2:
                # the selector looks at the normalised fused latents in x0's dtype (:1397)
                fused_lat = lib.latent_norm_replace(enc, x0, mean_h, inv_std_h, [])
                if self._selector is None:
                    sh = getattr(vae, "shard", None)      # one process per GPU: the ranks share the scoring work too
                    self._selector = flf_select.FlowChannelSelector() if sh is None or sh.world == 1 else \
                        flf_select.FlowChannelSelector(group=sh.group, world=sh.world, rank=sh.rank)
                with lib.phase("flf.channel_scoring"):
                    chans = self._selector.select(x0, fused_lat, step)
            self.flf_log.append((step, list(chans)))
        return lib.latent_norm_replace(enc, x0, mean_h, inv_std_h, chans)

    # ---------------------------------------------------------------------------------- step (:1457-1536)
    def step(self, model_output, timestep, sample, return_dict: bool = True, mask=None, guided: bool = False,
             video_latents=None, resampling: bool = False, vae=None, current_step: int = -1, resample_count: int = 2,
             is_resample_round: bool = False, static: bool = False, **kw):
        if self.num_inference_steps is None:
            raise ValueError("Run 'set_timesteps' after creating scheduler")
        if self._step_index is None:
            self._step_index = self._index_for_timestep(timestep) if self._begin_index is None else self._begin_index
        use_corrector = (self._step_index > 0 and self._step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        x0 = self.convert_model_output(model_output, sample=sample)
        if guided and video_latents is not None:
            x0 = self.fuse_latents(x0, video_latents, mask, vae=vae, current_step=current_step,
                                   total_steps=self.num_inference_steps, resampling=resampling, static=static, **kw)
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
        return SchedulerStepOutput(prev, x0)
