"""Oracle: the Wan 3D-VAE evaluated the way the reference EXECUTES it - chunk by chunk with per-convolution feature caches.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

``oracle/wan_vae.py`` states every layer over the whole clip (one function call per layer).  The reference instead walks
the clip in chunks - encode: frame 0 alone, then 4 frames at a time (vae.py:516-535); decode: one latent frame at a time
(:553-568) - and every causal convolution keeps the last ``CACHE_T = 2`` frames of its INPUT from the previous chunk
(CausalConv3d.forward :28-36; cache bookkeeping in ResidualBlock.forward :205-217, Encoder3d/Decoder3d.forward, and the
'Rep' / first-chunk special cases of Resample.forward :103-160).  This file restates that execution schedule over the
same parameter dictionary and layer plans:

* it pins the whole-clip form against the chunked form (tests/test_oracle_pinning.py: same function, ~1e-6), and
* on a GPU (device-agnostic torch, cuDNN convolutions) it IS the reference's VAE path - 21 x 33 small cached
  convolutions per decode - for bench.py's ``gpu_reference`` record (the PyTorch + flash-attn comparator of BASELINE.md §3).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .wan_vae import VaeConfig, attn_block, conv2d_frames, decoder_plan, encoder_plan, rms_norm_c

CACHE_T = 2     # vae.py:14


class _Caches:
    """feat_cache / feat_idx of the reference: one slot per causal convolution, visited in call order every chunk."""

    def __init__(self):
        self.slots: List[Optional[object]] = []
        self.idx = 0

    def begin_chunk(self):
        self.idx = 0

    def take(self):
        if self.idx == len(self.slots):
            self.slots.append(None)
        i = self.idx
        self.idx += 1
        return i


def _causal_conv(x, w, b, cache=None, stride_t: int = 1, pad_t: Optional[int] = None):
    """CausalConv3d.forward (:28-36): the cached frames replace that many zero frames of the causal left padding."""
    kt, kh, kw = w.shape[2:]
    pad_t = (kt - 1) if pad_t is None else pad_t
    if cache is not None and pad_t > 0:
        x = torch.cat([cache, x], dim=1)
        pad_t -= cache.shape[1]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pad_t, 0))
    return F.conv3d(x.unsqueeze(0), w, b, stride=(stride_t, 1, 1))[0]


def _cached_conv(C: _Caches, x, w, b):
    """The call-site pattern of ResidualBlock / Encoder3d / Decoder3d (:205-217): remember the last two input frames
    (topped up with the previous cache's last frame when the chunk has only one), convolve with the previous cache."""
    i = C.take()
    prev = C.slots[i]
    keep = x[:, -CACHE_T:].clone()
    if keep.shape[1] < 2 and prev is not None:
        keep = torch.cat([prev[:, -1:], keep], dim=1)
    y = _causal_conv(x, w, b, prev)
    C.slots[i] = keep
    return y


def _res_block(P, name, x, C: _Caches):
    h = x
    if name + ".shortcut.weight" in P:
        h = _causal_conv(x, P[name + ".shortcut.weight"], P[name + ".shortcut.bias"])        # 1x1x1: no cache (:36)
    y = F.silu(rms_norm_c(x, P[name + ".residual.0.gamma"]))
    y = _cached_conv(C, y, P[name + ".residual.2.weight"], P[name + ".residual.2.bias"])
    y = F.silu(rms_norm_c(y, P[name + ".residual.3.gamma"]))
    y = _cached_conv(C, y, P[name + ".residual.6.weight"], P[name + ".residual.6.bias"])
    return y + h


def _upsample(P, name, x, temporal: bool, C: _Caches):
    """Resample.forward, upsample2d / upsample3d (:101-140)."""
    Cc, T, H, W = x.shape
    if temporal:
        i = C.take()
        prev = C.slots[i]
        if prev is None:
            C.slots[i] = "Rep"                                   # first chunk: the frame passes through (:106-108)
        else:
            keep = x[:, -CACHE_T:].clone()
            if keep.shape[1] < 2:
                keep = torch.cat([prev[:, -1:] if not isinstance(prev, str) else torch.zeros_like(keep), keep], dim=1)
            t = _causal_conv(x, P[name + ".time_conv.weight"], P[name + ".time_conv.bias"], None if isinstance(prev, str) else prev)
            C.slots[i] = keep
            x = t.reshape(2, Cc, T, H, W).permute(1, 2, 0, 3, 4).reshape(Cc, 2 * T, H, W)
    u = F.interpolate(x.transpose(0, 1), scale_factor=(2.0, 2.0), mode="nearest-exact")
    y = F.conv2d(u, P[name + ".resample.1.weight"], P[name + ".resample.1.bias"], padding=1)
    return y.transpose(0, 1)


def _downsample(P, name, x, temporal: bool, C: _Caches):
    """Resample.forward, downsample2d / downsample3d (:141-159)."""
    y = conv2d_frames(x, P[name + ".resample.1.weight"], P[name + ".resample.1.bias"], stride=2, pad=(0, 1, 0, 1))
    if temporal:
        i = C.take()
        prev = C.slots[i]
        if prev is None:
            C.slots[i] = y.clone()                               # first chunk: no temporal convolution (:146-148)
        else:
            keep = y[:, -1:].clone()
            y = _causal_conv(torch.cat([prev[:, -1:], y], dim=1), P[name + ".time_conv.weight"], P[name + ".time_conv.bias"],
                             stride_t=2, pad_t=0)
            C.slots[i] = keep
    return y


def _run_chunk(P, plan, x, C: _Caches):
    C.begin_chunk()
    for kind, name, cin, cout in plan:
        if kind == "conv":
            x = _cached_conv(C, x, P[name + ".weight"], P[name + ".bias"])
        elif kind == "res":
            x = _res_block(P, name, x, C)
        elif kind == "attn":
            x = attn_block(P, name, x)
        elif kind in ("down2d", "down3d"):
            x = _downsample(P, name, x, kind == "down3d", C)
        elif kind in ("up2d", "up3d"):
            x = _upsample(P, name, x, kind == "up3d", C)
        elif kind == "head":
            x = F.silu(rms_norm_c(x, P[name + ".0.gamma"]))
            x = _cached_conv(C, x, P[name + ".2.weight"], P[name + ".2.bias"])
    return x


def encode_mode(P: Dict[str, torch.Tensor], cfg: VaeConfig, video: torch.Tensor) -> torch.Tensor:
    """WanVAE_.encode (:516-535) without the latent normalisation: chunks [1, 4, 4, ...], then conv1; returns the mean."""
    video = video.to(torch.float32)
    plan, C = encoder_plan(cfg), _Caches()
    outs = [_run_chunk(P, plan, video[:, :1], C)]
    for i in range(1, 1 + (video.shape[1] - 1) // 4):
        outs.append(_run_chunk(P, plan, video[:, 1 + 4 * (i - 1):1 + 4 * i], C))
    h = _causal_conv(torch.cat(outs, dim=1), P["conv1.weight"], P["conv1.bias"])
    return h[:cfg.z_dim]


def decode(P: Dict[str, torch.Tensor], cfg: VaeConfig, z: torch.Tensor) -> torch.Tensor:
    """WanVAE_.decode (:553-568) without the latent de-normalisation: conv2, then one latent frame per pass; clamped."""
    x = _causal_conv(z.to(torch.float32), P["conv2.weight"], P["conv2.bias"])
    plan, C = decoder_plan(cfg), _Caches()
    outs = [_run_chunk(P, plan, x[:, i:i + 1], C) for i in range(x.shape[1])]
    return torch.cat(outs, dim=1).clamp(-1.0, 1.0)
