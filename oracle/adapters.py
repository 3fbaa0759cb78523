"""Oracle objects with the duck-typed surfaces the reference pipeline calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``OracleTransformer`` quacks like ``pipe.transformer`` (call signature at
pipeline_wan_i2v_clean.py:593-610), ``OracleVAE`` like ``pipe.vae``
(``encode(x).latent_dist.mode()`` :348 / scheduling...:1384,
``decode(z, return_dict=False)[0]`` :743 / :1285, ``config.{z_dim,latents_mean,
latents_std}`` :333-340).  Batch size is 1, as in the entry script.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import wan_dit, wan_vae


class OracleTransformer:
    def __init__(self, params, cfg: wan_dit.DitConfig, amp: bool = True, out_dtype=torch.bfloat16):
        self.P, self.cfg, self.amp = params, cfg, amp
        self.dtype = out_dtype
        self.config = SimpleNamespace(patch_size=cfg.patch)
        self.calls = 0

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None,
                 attention_kwargs=None, return_dict=False):
        assert hidden_states.shape[0] == 1
        self.calls += 1
        y = wan_dit.dit_forward(self.P, self.cfg, hidden_states[0].cpu(), timestep.reshape(-1)[:1].cpu(),
                                encoder_hidden_states[0].cpu(), encoder_hidden_states_image[0].cpu(), amp=self.amp)
        return (y.unsqueeze(0).to(self.dtype).to(hidden_states.device),)


class _Dist:
    def __init__(self, mean):
        self._m = mean

    def mode(self):
        return self._m


class OracleVAE:
    """``device=None``: evaluated on the CPU in fp32.  ``device='cuda'``: the same oracle functions evaluated on that
    device - with ``torch.backends.cudnn.allow_tf32`` at its default (True) this is the arithmetic the reference's fp32
    VAE has on a GPU (cuDNN tf32 convolutions), with it False it is exact fp32."""

    def __init__(self, params, cfg: wan_vae.VaeConfig = wan_vae.WAN_VAE, device=None):
        self.dev = torch.device("cpu") if device is None else torch.device(device)
        self.P, self.cfg = {k: v.to(self.dev) for k, v in params.items()}, cfg
        self.dtype = torch.float32
        self.config = SimpleNamespace(z_dim=cfg.z_dim, latents_mean=list(wan_vae.LATENTS_MEAN[:cfg.z_dim]),
                                      latents_std=list(wan_vae.LATENTS_STD[:cfg.z_dim]))
        self.temperal_downsample = list(cfg.temporal_downsample)

    def encode(self, x):
        assert x.shape[0] == 1
        mu = wan_vae.encode_mode(self.P, self.cfg, x[0].to(self.dev))
        return SimpleNamespace(latent_dist=_Dist(mu.unsqueeze(0).to(x.device)))

    def decode(self, z, return_dict=False):
        assert z.shape[0] == 1
        return (wan_vae.decode(self.P, self.cfg, z[0].to(self.dev)).unsqueeze(0).to(z.device),)


class OracleLongCatDit:
    """Quacks like ``pipe.dit`` (pipeline_longcat_video.py:867-873) on top of oracle.longcat_dit."""

    def __init__(self, params, cfg, amp: bool = True, bsa=None):
        from . import longcat_dit
        self._ld = longcat_dit
        self.P, self.cfg, self.amp, self.bsa = params, cfg, amp, bsa
        self.dtype = torch.bfloat16
        self.config = SimpleNamespace(in_channels=cfg.in_channels)
        self.cp_split_hw = [1, 1]
        self.calls = 0

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask=None, num_cond_latents=0,
                 return_kv=False, kv_cache_dict=None, skip_crs_attn=False, **kw):
        outs, caches = [], []
        for s in range(hidden_states.shape[0]):
            ctx = encoder_hidden_states[s, 0]
            if encoder_attention_mask is not None and not skip_crs_attn:
                ctx = ctx[encoder_attention_mask[s].reshape(-1) != 0]
            self.calls += 1
            cache_s = None
            if kv_cache_dict:                                   # [B, N, H, D] per layer; batch 1 is shared (attention.py:163-166)
                cache_s = {i: (kv[0][min(s, kv[0].shape[0] - 1)], kv[1][min(s, kv[1].shape[0] - 1)]) for i, kv in kv_cache_dict.items()}
            r = self._ld.dit_forward(self.P, self.cfg, hidden_states[s].cpu(), timestep[s].cpu(), ctx.cpu(),
                                     num_cond_latents=num_cond_latents, amp=self.amp, bsa=self.bsa, return_kv=return_kv,
                                     kv_cache_dict=cache_s, skip_crs_attn=skip_crs_attn)
            if return_kv:
                outs.append(r[0]); caches.append(r[1])
            else:
                outs.append(r)
        out = torch.stack(outs).to(hidden_states.device)
        if not return_kv:
            return out
        return out, {i: (torch.stack([c[i][0] for c in caches]), torch.stack([c[i][1] for c in caches])) for i in caches[0]}
