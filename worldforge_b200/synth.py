"""Synthetic warped inputs for the WorldForge sampling path (no checkpoints, no datasets).

Host-side input preparation, shaped as SURVEY.md §8(d) prescribes: a smooth random
RGB field translated by 2 px per frame stands in for the warped reference clip
(so Farneback sees real motion), a moving half-plane softened on its inside edge
stands in for the validity mask, and N(0,1) tensors stand in for the T5 / CLIP
embeddings.  ``soften_mask`` has the semantics of the entry script's helper
(reference infer_worldforge.py:105-150: Euclidean distance transform of the valid
region, sine ramp over ``transition_distance`` pixels).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


def soften_mask(mask: np.ndarray, transition_distance: int = 15, decay_type: str = "sine") -> np.ndarray:
    """[F,H,W] {0,1} -> float32 with a smooth 0->1 ramp just inside each valid region."""
    from scipy.ndimage import distance_transform_edt
    ramps = {
        "linear": lambda t: t,
        "exponential": lambda t: 1.0 - np.exp(-3.0 * t),
        "sine": lambda t: np.sin(np.pi / 2 * t),
        "cosine": lambda t: 1.0 - np.cos(np.pi / 2 * t),
    }
    if decay_type not in ramps:
        raise ValueError(f"Unsupported decay type: {decay_type}")
    out = mask.astype(np.float32).copy()
    for f in range(mask.shape[0]):
        valid = mask[f].astype(bool)
        if valid.all() or (~valid).all():
            continue
        dist = distance_transform_edt(valid)
        edge = valid & (dist <= transition_distance)
        if edge.any():
            out[f][edge] = ramps[decay_type](np.clip(dist[edge] / transition_distance, 0.0, 1.0))
    return out


def smooth_video(num_frames: int, height: int, width: int, seed: int = 42, shift_px: int = 2) -> torch.Tensor:
    """[1,3,F,H,W] in [0,1]: one smooth random image, rolled by ``shift_px`` per frame."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(1, 3, max(height // 16, 2), max(width // 16, 2), generator=g)
    img = torch.nn.functional.interpolate(coarse, size=(height, width), mode="bicubic", align_corners=False)
    img = img.clamp(0.0, 1.0)[0]
    frames = [torch.roll(img, shifts=shift_px * f, dims=2) for f in range(num_frames)]
    return torch.stack(frames, dim=1).unsqueeze(0).contiguous()


def moving_mask(num_frames: int, height: int, width: int, soften: bool = True) -> torch.Tensor:
    """[1,1,F,H,W]: frame 0 fully valid, then a valid half-plane whose edge moves right."""
    m = np.zeros((num_frames, height, width), dtype=np.float32)
    m[0] = 1.0
    for f in range(1, num_frames):
        edge = int(width * (0.35 + 0.4 * f / max(num_frames - 1, 1)))
        m[f, :, :edge] = 1.0
    if soften:
        m = soften_mask(m, 15, "sine")
    return torch.from_numpy(m).unsqueeze(0).unsqueeze(0)


@dataclass
class SynthInputs:
    latents: torch.Tensor            # [1,16,f,h,w] fp32 initial noise
    condition: torch.Tensor          # [1,20,f,h,w] fp32 (4 mask channels + 16 latent channels)
    prompt_embeds: torch.Tensor      # [1,text_len,text_dim] bf16
    negative_prompt_embeds: torch.Tensor
    image_embeds: torch.Tensor       # [1,img_len,img_dim] bf16
    video_ref: torch.Tensor          # [1,3,F,H,W] fp32 in [0,1]
    mask: torch.Tensor               # [1,1,F,H,W] fp32 in [0,1]


def frame_mask_channels(num_frames: int, lat_h: int, lat_w: int, t_scale: int = 4) -> torch.Tensor:
    """The 4 first-frame mask channels of the I2V condition
    (reference pipeline_wan_i2v_clean.py:353-360): [1,4,f,h,w]."""
    m = torch.zeros(1, 1, num_frames + t_scale - 1, lat_h, lat_w)
    m[:, :, :t_scale] = 1.0
    return m.view(1, -1, t_scale, lat_h, lat_w).transpose(1, 2).contiguous()


def make_inputs(num_frames: int, height: int, width: int, text_len: int = 512, text_dim: int = 4096,
                img_len: int = 257, img_dim: int = 1280, z_dim: int = 16, seed: int = 42,
                real_tokens: int = 64) -> SynthInputs:
    g = torch.Generator().manual_seed(seed)
    f, h, w = (num_frames - 1) // 4 + 1, height // 8, width // 8
    latents = torch.randn(1, z_dim, f, h, w, generator=g)
    cond_lat = torch.randn(1, z_dim, f, h, w, generator=g)
    condition = torch.cat([frame_mask_channels(num_frames, h, w), cond_lat], dim=1)
    def text():
        e = torch.randn(1, text_len, text_dim, generator=g)
        e[:, min(real_tokens, text_len):] = 0          # zero padding after the real tokens
        return e.to(torch.bfloat16)
    pe, ne = text(), text()
    ie = torch.randn(1, img_len, img_dim, generator=g).to(torch.bfloat16)
    return SynthInputs(latents, condition, pe, ne, ie,
                       smooth_video(num_frames, height, width, seed),
                       moving_mask(num_frames, height, width))
