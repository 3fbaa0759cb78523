"""Input preparation of the entry script on the device.

Stands where ``wan_for_worldforge/infer_worldforge.py`` prepares the warped reference clip and its mask: ``soften_mask``
(:105-150, same name and arguments), the frame stacking (:232-238) and the mask stacking (:244-251).  PIL's resize of the
PNG frames (:225, :231, :245) stays on the host - it happens once per video on 8-bit images.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib


def _ramp_table(transition_distance, decay_type: str):
    """smooth(sqrt(d2) / transition_distance) for every integer d2 <= transition_distance^2, evaluated with the reference's
    float64 numpy expressions (:126-137, :141-143) and stored to float32 as the reference's assignment does (:143)."""
    td = float(transition_distance)
    max_d2 = int(np.floor(td * td))
    d = np.sqrt(np.arange(max_d2 + 1, dtype=np.float64))
    t = np.clip(d / transition_distance, 0.0, 1.0)
    if decay_type == "linear":
        v = t
    elif decay_type == "exponential":
        v = 1.0 - np.exp(-3.0 * t)
    elif decay_type == "sine":
        v = np.sin(np.pi / 2 * t)
    elif decay_type == "cosine":
        v = 1.0 - np.cos(np.pi / 2 * t)
    else:
        raise ValueError(f"Unsupported decay type: {decay_type}")
    return v.astype(np.float32), int(np.floor(td)), max_d2


_U8_TO_UNIT = torch.from_numpy((np.arange(256) / 255.0).astype(np.float32))     # float32(u8 / 255.0 in float64), :246 then :117


def mask_from_u8(masks_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [F,H,W] on the device -> fp32 [F,H,W] = float32(np.array(mask) / 255.0)."""
    return _U8_TO_UNIT.to(masks_u8.device)[masks_u8.long()]


def soften_mask(mask_array, transition_distance=15, decay_type: str = "sine") -> torch.Tensor:
    """``mask_array`` [F,H,W]: CUDA tensor (fp32 values 0..1, or uint8 0..255) or a numpy array (uploaded) -> fp32 CUDA tensor,
    the reference's softened mask bit for bit."""
    if isinstance(mask_array, np.ndarray):
        mask_array = torch.from_numpy(np.ascontiguousarray(mask_array)).to("cuda")
    if not mask_array.is_cuda:
        raise lib.WfError("soften_mask runs on CUDA tensors only (no CPU fallback)")
    m = mask_from_u8(mask_array) if mask_array.dtype == torch.uint8 else mask_array.to(torch.float32)
    m = m.contiguous()
    assert m.dim() == 3
    if transition_distance > 31:
        raise lib.WfError("soften_mask: transition_distance above 31 pixels is not supported on the device")
    table, radius, max_d2 = _ramp_table(transition_distance, decay_type)
    lut = torch.from_numpy(table).to(m.device)
    out = torch.empty_like(m)
    lib._call("wf_soften_mask", m.data_ptr(), out.data_ptr(), m.shape[0], m.shape[1], m.shape[2], radius, max_d2, lut.data_ptr(),
              lib._stream())
    return out


def clip_from_frames(frames_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [F,H,W,3] on the device -> video_ref [1,3,F,H,W] fp32 in [0,1] (:232-238)."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise lib.WfError("clip_from_frames: expected a CUDA uint8 tensor [F,H,W,3]")
    f = frames_u8.contiguous()
    F_, H, W = f.shape[:3]
    out = torch.empty(1, 3, F_, H, W, dtype=torch.float32, device=f.device)
    lib._call("wf_clip_from_u8", f.data_ptr(), out.data_ptr(), F_ * H * W, lib._stream())
    return out


def prepare_mask(masks_u8: torch.Tensor, soften: bool = True, transition_distance=15, decay_type: str = "sine") -> torch.Tensor:
    """uint8 [F,H,W] -> mask [1,1,F,H,W] fp32 (:244-251)."""
    m = soften_mask(masks_u8, transition_distance, decay_type) if soften else mask_from_u8(masks_u8)
    return m.unsqueeze(0).unsqueeze(0)


def presize_guidance(video_ref: torch.Tensor, mask: torch.Tensor, target_shape):
    """The resize branch of ``fuse_latents`` (utils/scheduling_unipc_multistep_clean.py:1297-1371): a warped clip
    [b,3,F,h,w] / mask [b,c,F,h,w] whose size is not the decoded clip's is brought to it - batch repeated, the clip
    resampled in the plane with bilinear taps (align_corners=False), the mask cut to its first channel and resampled with
    nearest taps.  Runs once per clip (the scheduler keeps the result), on whatever device the tensors live on, with
    torch's own resampling - the op the reference calls, so the values are the reference's.  As in the reference, a
    different frame count is an error (its temporal branch hands ``F.interpolate`` a 4-D tensor with a 3-element size,
    :1326-1334), and so is a mask that needs resampling next to a clip that does not (``F`` is bound at :1301 only)."""
    import torch.nn.functional as F
    B, C, T, H, W = (int(v) for v in target_shape)
    if video_ref.dim() != 5 or mask.dim() != 5:
        raise ValueError(f"video_ref / mask must be 5-D, got {tuple(video_ref.shape)} / {tuple(mask.shape)}")
    if tuple(video_ref.shape) == (B, C, T, H, W) and tuple(mask.shape[-2:]) != (H, W):
        raise ValueError(f"mask {tuple(mask.shape)} needs a spatial resize to {(H, W)} but video_ref does not: "
                         "the reference fails on this input (scheduling…:1301, :1356)")
    if video_ref.shape[2] != T or mask.shape[2] != T:
        raise ValueError(f"video_ref / mask have {video_ref.shape[2]} / {mask.shape[2]} frames, the decoded clip {T}: "
                         "the reference cannot resample in time (scheduling…:1326-1334, :1363-1370 raise)")
    if tuple(video_ref.shape) != (B, C, T, H, W):
        if video_ref.shape[0] != B:
            video_ref = video_ref.repeat(B, 1, 1, 1, 1)
        b, c, f, h, w = video_ref.shape
        if (h, w) != (H, W):
            video_ref = F.interpolate(video_ref.reshape(b * c * f, 1, h, w), size=(H, W), mode="bilinear",
                                      align_corners=False).reshape(b, c, f, H, W)
    if tuple(mask.shape) != (B, 1, T, H, W):
        if mask.shape[0] != B:
            mask = mask.repeat(B, 1, 1, 1, 1)
        if mask.shape[1] != 1:
            mask = mask[:, 0:1]
        b, c, f, h, w = mask.shape
        if (h, w) != (H, W):
            mask = F.interpolate(mask.reshape(b * c * f, 1, h, w), size=(H, W), mode="nearest").reshape(b, c, f, H, W)
    return video_ref.contiguous(), mask.contiguous()


# ---- the on-disk contract of the warping stages: a folder of warped frames and ``mask_*`` images -----------------------------
_IMAGE_PATTERNS = ("*.jpg", "*.jpeg", "*.png", "*.bmp", "*.tiff")


def list_case_folder(directory: str):
    """(frame files, mask files) of a warp folder in the reference's order (infer_worldforge.py:65-86;
    run_longcat_worldforge_single.py:56-79): every image file of the five extensions, sorted by full path, split by the
    ``mask_`` prefix of the file name."""
    import glob
    import os
    files = []
    for pat in _IMAGE_PATTERNS:
        files.extend(glob.glob(os.path.join(directory, pat)))
    files = sorted(files)
    if not files:
        raise ValueError(f"No image files found in directory {directory}")
    masks = [f for f in files if os.path.basename(f).startswith("mask_")]
    frames = [f for f in files if not os.path.basename(f).startswith("mask_")]
    return frames, masks


def read_case_folder(directory: str):
    """``read_frames_from_directory`` (infer_worldforge.py:65-102): (RGB frames, L masks, first frame) as PIL images; no mask
    files -> all-zero masks, fewer masks than frames -> the last one repeated, more -> cut (:92-99)."""
    from PIL import Image
    frame_files, mask_files = list_case_folder(directory)
    frames = [Image.open(f).convert("RGB") for f in frame_files]
    masks = [Image.open(f).convert("L") for f in mask_files]
    if not masks and frames:
        masks = [Image.new("L", frames[0].size, 0) for _ in frames]
    while len(masks) < len(frames):
        masks.append(masks[-1] if masks else Image.new("L", frames[0].size, 0))
    masks = masks[:len(frames)]
    return frames, masks, (frames[0] if frames else None)


def target_size(image_width: int, image_height: int, max_area: int = 480 * 832, mod_value: int = 16):
    """(width, height) the entry script resizes everything to (:218-221): the area budget at the first frame's aspect
    ratio, rounded, then cut to a multiple of vae_scale_factor_spatial * patch_size = 16."""
    aspect_ratio = image_height / image_width
    height = round(np.sqrt(max_area * aspect_ratio)) // mod_value * mod_value
    width = round(np.sqrt(max_area / aspect_ratio)) // mod_value * mod_value
    return int(width), int(height)


def load_case(directory: str, max_area: int = 480 * 832, mod_value: int = 16, soften: bool = True, transition_distance=15,
              decay_type: str = "sine", device="cuda", image=None, _to_clip=None, _to_mask=None):
    """A warp folder -> what the pipeline call takes (infer_worldforge.py:156, :208-251): ``image`` (PIL, resized),
    ``video_ref`` [1,3,F,H,W] fp32 in [0,1] and ``mask`` [1,1,F,H,W] fp32 on ``device``, plus ``width`` / ``height``.
    PIL decodes and resizes the 8-bit images on the host exactly as the script does (its default filter); the uint8 stacks
    are uploaded once and turned into the fp32 clip / softened mask by the device kernels above.
    (``_to_clip`` / ``_to_mask`` replace those two device steps - the CPU tests check the composition with the oracle's.)"""
    frames, masks, first = read_case_folder(directory)
    image = first if image is None else image
    if image is None:
        raise ValueError("Cannot get first frame as input image, please specify an image")
    width, height = target_size(image.width, image.height, max_area, mod_value)
    image = image.resize((width, height))
    frames_u8 = torch.from_numpy(np.stack([np.array(f.resize((width, height))) for f in frames]))          # [F,H,W,3]
    masks_u8 = torch.from_numpy(np.stack([np.array(m.resize((width, height))) for m in masks]))            # [F,H,W]
    to_clip = _to_clip if _to_clip is not None else (lambda u8: clip_from_frames(u8.to(device)))
    to_mask = _to_mask if _to_mask is not None else \
        (lambda u8: prepare_mask(u8.to(device), soften=soften, transition_distance=transition_distance, decay_type=decay_type))
    return dict(image=image, video_ref=to_clip(frames_u8), mask=to_mask(masks_u8), width=width, height=height,
                num_frames=len(frames))
