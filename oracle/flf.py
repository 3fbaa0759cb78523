"""Oracle: FLF (flow-gated latent fusion) channel scoring and selection.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``VideoMotionPCASelector`` in
``wan_for_worldforge/utils/scheduling_unipc_multistep_clean.py``:
select_motion_related_channels :338-437, _extract_optical_flow_motion :156-248,
_compute_channel_correlations :439-495, _compute_flow_metrics :497-607.
Dense optical flow is OpenCV's ``calcOpticalFlowFarneback`` exactly as the
reference calls it (:220-224) - OpenCV is the reference's own third-party
dependency for this step (requirements.txt:8, version unpinned; 4.13.0 in this
image), not something this repo re-implements.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

FARNEBACK = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def quantise_u8(lat: torch.Tensor) -> np.ndarray:
    """[1,C,T,H,W] latents -> uint8 [C,T,H,W]: min-max normalised with the GLOBAL
    min/max of the tensor (:376-378, :462-464), times 255, truncated (:175-176)."""
    f = lat.to(torch.float32)
    lo, hi = f.min(), f.max()
    rng = hi - lo + 1e-8
    n = ((f - lo) / rng)[0].cpu().numpy()
    return (n * 255).astype(np.uint8)


def farneback_flows(u8: np.ndarray) -> torch.Tensor:
    """uint8 [T,H,W] -> flow [1, T-1, 2, H, W] fp32 (:193-248).  The reference
    replicates the channel to RGB and converts back to gray (:200-201)."""
    import cv2
    flows = []
    for t in range(u8.shape[0] - 1):
        a = cv2.cvtColor(np.repeat(u8[t][:, :, None], 3, axis=2), cv2.COLOR_RGB2GRAY)
        b = cv2.cvtColor(np.repeat(u8[t + 1][:, :, None], 3, axis=2), cv2.COLOR_RGB2GRAY)
        flows.append(cv2.calcOpticalFlowFarneback(a, b, None, **FARNEBACK))
    fl = np.stack(flows, axis=0).transpose(0, 3, 1, 2)
    t = torch.from_numpy(fl).float().unsqueeze(0)
    if not torch.isfinite(t).all():
        ok = torch.isfinite(t)
        t = torch.where(ok, t, t[ok].mean() if ok.any() else torch.zeros(()))
    return t


def flow_similarity(ref_flow: torch.Tensor, chan_flow: torch.Tensor) -> float:
    """1 - weighted(M-EPE, Fl-all, M-AE) (:541-604); inputs [1, T-1, 2, H, W]."""
    d = ref_flow - chan_flow
    epe = torch.sqrt((d ** 2).sum(dim=2) + 1e-8)
    dot = (ref_flow * chan_flow).sum(dim=2)
    rn = torch.sqrt((ref_flow ** 2).sum(dim=2) + 1e-8)
    cn = torch.sqrt((chan_flow ** 2).sum(dim=2) + 1e-8)
    cos = torch.clamp(dot / (rn * cn + 1e-8), -1.0, 1.0)
    ang = torch.acos(cos) * 180.0 / torch.pi
    outlier = (epe > 3.0) & (epe > rn * 0.05)
    n_epe = torch.clamp(epe.mean() / 10.0, 0.0, 1.0)
    n_fl = torch.clamp(outlier.float().mean() / 0.5, 0.0, 1.0)
    n_ae = torch.clamp(ang.mean() / 30.0, 0.0, 1.0)
    err = 0.45 * n_epe + 0.45 * n_fl + 0.1 * n_ae
    return torch.clamp(1.0 - err, 0.0, 1.0).item()


def channel_scores(pred_x0: torch.Tensor, fused: torch.Tensor) -> List[float]:
    """Per-channel similarity between the flow of the fused (reference-carrying)
    latents and the flow of the model's own prediction (:380-403, :466-495)."""
    ref_u8 = quantise_u8(fused)
    pred_u8 = quantise_u8(pred_x0)
    return [flow_similarity(farneback_flows(ref_u8[c]), farneback_flows(pred_u8[c]))
            for c in range(pred_x0.shape[1])]


def policy(scores: List[float], step: int) -> List[int]:
    """Which channels keep the model's own prediction (:408-437)."""
    s = np.array(scores)
    if step <= 10:
        take = 0 if step <= 5 else 1
        out = np.argsort(s)[:take].tolist()
    else:
        thr = np.mean(s) - 0.625 * np.std(s)
        below = [i for i, v in enumerate(s) if v < thr]
        if len(below) < 2:
            out = np.argsort(s)[:2].tolist()
        elif len(below) > 6:
            out = [i for i, _ in sorted(((i, s[i]) for i in below), key=lambda p: p[1])[:6]]
        else:
            out = below
    return sorted(out)


def select_channels(pred_x0: torch.Tensor, fused: torch.Tensor, step: int) -> List[int]:
    if step < 2:                       # :364
        return []
    return policy(channel_scores(pred_x0, fused), step)


def presize_guidance(video_ref: torch.Tensor, mask: torch.Tensor, target_shape):
    """Bring the warped clip [b,3,F,h,w] and its mask [b,c,F,h,w] to the decoded clip's shape (:1297-1371): batch
    repeated, the clip resized in the plane with bilinear taps (align_corners=False, every frame of every channel on its
    own, :1316-1324), the mask reduced to its first channel (:1342-1343) and resized with nearest taps (:1354-1361).  A
    frame-count mismatch reaches ``F.interpolate`` with a 4-D tensor and a 3-element size (:1326-1334, :1363-1370), which
    torch rejects with a ValueError - the reference cannot resample in time either, so neither does this."""
    import torch.nn.functional as F
    B, C, T, H, W = (int(v) for v in target_shape)
    if tuple(video_ref.shape) == (B, C, T, H, W) and tuple(mask.shape[-2:]) != (H, W):
        # the reference imports F inside the clip branch only (:1301): a mask that needs resizing next to a clip that does not
        # dies with UnboundLocalError at :1356
        raise ValueError("mask needs a spatial resize but video_ref does not: the reference fails here (F is bound at :1301 only)")
    if tuple(video_ref.shape) != (B, C, T, H, W):
        if video_ref.shape[0] != B:
            video_ref = video_ref.repeat(B, 1, 1, 1, 1)
        b, c, f, h, w = video_ref.shape
        if (h, w) != (H, W):
            video_ref = F.interpolate(video_ref.reshape(b * c * f, h, w).unsqueeze(1), size=(H, W), mode="bilinear",
                                      align_corners=False).squeeze(1).reshape(b, c, f, H, W)
        if f != T:
            raise ValueError(f"video_ref has {f} frames, the decoded clip {T}: the reference's temporal branch "
                             "(:1326-1334) calls F.interpolate with a 4-D input and a 3-D size and raises")
    if tuple(mask.shape) != (B, 1, T, H, W):
        if mask.shape[0] != B:
            mask = mask.repeat(B, 1, 1, 1, 1)
        if mask.shape[1] != 1:
            mask = mask[:, 0:1]
        b, c, f, h, w = mask.shape
        if (h, w) != (H, W):
            mask = F.interpolate(mask.reshape(b * c * f, h, w).unsqueeze(1), size=(H, W), mode="nearest").squeeze(1) \
                .reshape(b, c, f, H, W)
        if f != T:
            raise ValueError(f"mask has {f} frames, the decoded clip {T}: the reference's temporal branch (:1363-1370) raises")
    return video_ref, mask
