"""Oracle: Wan2.1 causal 3D-VAE encode / decode, restated on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``wan_for_worldforge/wan/modules/vae.py`` (CausalConv3d :17-36,
RMS_norm :39-54, Resample :66-160, ResidualBlock :186-220, AttentionBlock
:223-262, Encoder3d :265-366, Decoder3d :369-472, WanVAE_.encode/decode
:516-568), which is the same network as the ``AutoencoderKLWan`` the entry
script loads (infer_worldforge.py:185-189).

The reference walks the clip in chunks (1 frame, then 4 at a time when
encoding; 1 latent frame at a time when decoding) and carries the last two
input frames of every causal convolution in a feature cache.  This restatement
evaluates each layer over the WHOLE clip instead; the two are the same function
because of three facts read off the reference and checked numerically against
it (tests/test_oracle_pinning.py):

* a cached causal 3x3x3 convolution is a convolution with two zero frames of
  left padding in time (:28-36 with the cache logic :205-217);
* ``downsample3d`` passes the first frame through and convolves the rest with
  stride 2 from frame 0 (:143-159);
* ``upsample3d`` passes the first frame through and feeds frames 1.. to the
  temporal convolution with an all-zero history (:103-137), de-interleaving
  its 2C output channels into two consecutive frames.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

F32 = torch.float32

LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]   # vae.py:629-632
LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
               3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]        # vae.py:633-636


@dataclass
class VaeConfig:
    dim: int = 96
    z_dim: int = 16
    dim_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temporal_downsample: Tuple[bool, ...] = (False, True, True)   # vae.py:603


WAN_VAE = VaeConfig()


# ---- architecture plan ----------------------------------------------------

def encoder_plan(cfg: VaeConfig) -> List[tuple]:
    dims = [cfg.dim * u for u in (1,) + tuple(cfg.dim_mult)]
    plan = [("conv", "encoder.conv1", 3, dims[0])]
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg.num_res_blocks):
            plan.append(("res", f"encoder.downsamples.{idx}", cin, cout)); idx += 1
            cin = cout
        if i != len(cfg.dim_mult) - 1:
            mode = "down3d" if cfg.temporal_downsample[i] else "down2d"
            plan.append((mode, f"encoder.downsamples.{idx}", cout, cout)); idx += 1
    c = dims[-1]
    plan += [("res", "encoder.middle.0", c, c), ("attn", "encoder.middle.1", c, c),
             ("res", "encoder.middle.2", c, c),
             ("head", "encoder.head", c, cfg.z_dim * 2)]
    return plan


def decoder_plan(cfg: VaeConfig) -> List[tuple]:
    dims = [cfg.dim * u for u in (cfg.dim_mult[-1],) + tuple(cfg.dim_mult[::-1])]
    up = tuple(cfg.temporal_downsample[::-1])
    c = dims[0]
    plan = [("conv", "decoder.conv1", cfg.z_dim, c),
            ("res", "decoder.middle.0", c, c), ("attn", "decoder.middle.1", c, c),
            ("res", "decoder.middle.2", c, c)]
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg.num_res_blocks + 1):
            plan.append(("res", f"decoder.upsamples.{idx}", cin, cout)); idx += 1
            cin = cout
        if i != len(cfg.dim_mult) - 1:
            mode = "up3d" if up[i] else "up2d"
            plan.append((mode, f"decoder.upsamples.{idx}", cout, cout // 2)); idx += 1
    plan.append(("head", "decoder.head", dims[-1], 3))
    return plan


def param_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    def conv3(name, cin, cout, k):
        s[name + ".weight"] = (cout, cin) + k; s[name + ".bias"] = (cout,)
    def res(name, cin, cout):
        s[name + ".residual.0.gamma"] = (cin, 1, 1, 1)
        conv3(name + ".residual.2", cin, cout, (3, 3, 3))
        s[name + ".residual.3.gamma"] = (cout, 1, 1, 1)
        conv3(name + ".residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv3(name + ".shortcut", cin, cout, (1, 1, 1))
    for plan in (encoder_plan(cfg), decoder_plan(cfg)):
        for kind, name, cin, cout in plan:
            if kind == "conv":
                conv3(name, cin, cout, (3, 3, 3))
            elif kind == "res":
                res(name, cin, cout)
            elif kind == "attn":
                s[name + ".norm.gamma"] = (cin, 1, 1)
                s[name + ".to_qkv.weight"] = (3 * cin, cin, 1, 1); s[name + ".to_qkv.bias"] = (3 * cin,)
                s[name + ".proj.weight"] = (cin, cin, 1, 1); s[name + ".proj.bias"] = (cin,)
            elif kind in ("down2d", "down3d"):
                s[name + ".resample.1.weight"] = (cin, cin, 3, 3); s[name + ".resample.1.bias"] = (cin,)
                if kind == "down3d":
                    conv3(name + ".time_conv", cin, cin, (3, 1, 1))
            elif kind in ("up2d", "up3d"):
                s[name + ".resample.1.weight"] = (cin // 2, cin, 3, 3); s[name + ".resample.1.bias"] = (cin // 2,)
                if kind == "up3d":
                    conv3(name + ".time_conv", cin, 2 * cin, (3, 1, 1))
            elif kind == "head":
                s[name + ".0.gamma"] = (cin, 1, 1, 1)
                conv3(name + ".2", cin, cout, (3, 3, 3))
    conv3("conv1", cfg.z_dim * 2, cfg.z_dim * 2, (1, 1, 1))
    conv3("conv2", cfg.z_dim, cfg.z_dim, (1, 1, 1))
    return s


def init_params(cfg: VaeConfig, seed: int = 4321, device="cpu") -> Dict[str, torch.Tensor]:
    """Random-init fp32 weights: N(0, 1/fan_in) kernels, gains 1+N(0,0.05^2),
    biases N(0, 0.02^2)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("gamma"):
            w = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            w = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) / fan_in ** 0.5
        out[name] = w.to(device)
    return out


# ---- layers (x is [C, T, H, W], one clip) ----------------------------------

def causal_conv3d(x, w, b, stride_t: int = 1, pad_t=None):
    kt, kh, kw = w.shape[2:]
    if pad_t is None:
        pad_t = kt - 1
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pad_t, 0))
    return F.conv3d(x.unsqueeze(0), w, b, stride=(stride_t, 1, 1))[0]


def rms_norm_c(x, gamma):
    """RMS_norm (vae.py:51-54): L2-normalise over channels, times sqrt(C)*gamma."""
    return F.normalize(x, dim=0) * (x.shape[0] ** 0.5) * gamma.reshape(-1, *([1] * (x.dim() - 1)))


def conv2d_frames(x, w, b, stride=1, pad=(1, 1, 1, 1)):
    """A 2-D conv applied to every frame of [C,T,H,W]."""
    y = F.conv2d(F.pad(x.transpose(0, 1), pad), w, b, stride=stride)
    return y.transpose(0, 1)


def res_block(P, name, x):
    h = x
    if name + ".shortcut.weight" in P:
        h = causal_conv3d(x, P[name + ".shortcut.weight"], P[name + ".shortcut.bias"])
    y = F.silu(rms_norm_c(x, P[name + ".residual.0.gamma"]))
    y = causal_conv3d(y, P[name + ".residual.2.weight"], P[name + ".residual.2.bias"])
    y = F.silu(rms_norm_c(y, P[name + ".residual.3.gamma"]))
    y = causal_conv3d(y, P[name + ".residual.6.weight"], P[name + ".residual.6.bias"])
    return y + h


def attn_block(P, name, x):
    C, T, H, W = x.shape
    f = x.transpose(0, 1)                                     # [T, C, H, W]
    n = F.normalize(f, dim=1) * (C ** 0.5) * P[name + ".norm.gamma"]
    qkv = F.conv2d(n, P[name + ".to_qkv.weight"], P[name + ".to_qkv.bias"])
    q, k, v = qkv.reshape(T, 3, C, H * W).permute(1, 0, 3, 2)  # each [T, HW, C]
    o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1))[:, 0]
    o = o.permute(0, 2, 1).reshape(T, C, H, W)
    o = F.conv2d(o, P[name + ".proj.weight"], P[name + ".proj.bias"])
    return (o + f).transpose(0, 1)


def downsample(P, name, x, temporal: bool):
    y = conv2d_frames(x, P[name + ".resample.1.weight"], P[name + ".resample.1.bias"],
                      stride=2, pad=(0, 1, 0, 1))
    if temporal and y.shape[1] > 1:
        t = causal_conv3d(y, P[name + ".time_conv.weight"], P[name + ".time_conv.bias"],
                          stride_t=2, pad_t=0)
        y = torch.cat([y[:, :1], t], dim=1)
    return y


def upsample(P, name, x, temporal: bool):
    C, T, H, W = x.shape
    if temporal and T > 1:
        t = causal_conv3d(x[:, 1:], P[name + ".time_conv.weight"], P[name + ".time_conv.bias"])
        t = t.reshape(2, C, T - 1, H, W).permute(1, 2, 0, 3, 4).reshape(C, 2 * (T - 1), H, W)
        x = torch.cat([x[:, :1], t], dim=1)
    u = F.interpolate(x.transpose(0, 1), scale_factor=(2.0, 2.0), mode="nearest-exact")
    y = F.conv2d(u, P[name + ".resample.1.weight"], P[name + ".resample.1.bias"], padding=1)
    return y.transpose(0, 1)


def _run(P, plan, x):
    for kind, name, cin, cout in plan:
        if kind == "conv":
            x = causal_conv3d(x, P[name + ".weight"], P[name + ".bias"])
        elif kind == "res":
            x = res_block(P, name, x)
        elif kind == "attn":
            x = attn_block(P, name, x)
        elif kind in ("down2d", "down3d"):
            x = downsample(P, name, x, kind == "down3d")
        elif kind in ("up2d", "up3d"):
            x = upsample(P, name, x, kind == "up3d")
        elif kind == "head":
            x = F.silu(rms_norm_c(x, P[name + ".0.gamma"]))
            x = causal_conv3d(x, P[name + ".2.weight"], P[name + ".2.bias"])
    return x


def encode_mode(P, cfg: VaeConfig, video):
    """video [3, F, H, W] in [-1,1] (F = 4k+1) -> latent mean [z, f, H/8, W/8]
    (un-normalised; ``vae.encode(x).latent_dist.mode()``, vae.py:516-535)."""
    h = _run(P, encoder_plan(cfg), video.to(F32))
    h = causal_conv3d(h, P["conv1.weight"], P["conv1.bias"])
    return h[:cfg.z_dim]


def decode(P, cfg: VaeConfig, z):
    """latent [z, f, h, w] (un-normalised) -> video [3, 4(f-1)+1, 8h, 8w] clamped to
    [-1, 1] (vae.py:553-568 and the clamp at :661 / in AutoencoderKLWan._decode)."""
    h = causal_conv3d(z.to(F32), P["conv2.weight"], P["conv2.bias"])
    return _run(P, decoder_plan(cfg), h).clamp(-1.0, 1.0)
