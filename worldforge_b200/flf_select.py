"""FLF channel selection: which latent channels keep the model's own prediction.

Host side of the reference's ``VideoMotionPCASelector`` (reference
utils/scheduling_unipc_multistep_clean.py:338-607).  The reference scores each of the 16
latent channels by comparing dense optical flow (OpenCV Farneback, :220-224) of the fused
latents with that of the model's prediction, after quantising each channel to uint8.  Here

* the min-max quantisation of both 16-channel tensors runs on the GPU (``wf_quantise_u8``,
  two passes over 2 MB) so 4 MB of uint8 cross PCIe instead of 16 MB of fp32 in 32 copies;
* Farneback - OpenCV in the reference, its own third-party dependency for this step, whose discrete
  outcome (an argsort over 16 scores) must not drift - runs on the GPU for frames whose pyramid has one
  level (the 60 x 104 latent frames of 480p) or two (90 x 160 at 720p): ``wf_farneback_u8`` +
  ``wf_flow_metrics``, 2.3 ms against 81 ms for a 16-channel x 21-frame call at 480p, flows equal to
  OpenCV's to ~3e-6 px (one level) / ~1e-4 px (two levels), identical scores and selections
  (tests/test_flow_gpu.py).  Other frame sizes and WF_FLF_GPU=0 keep OpenCV on
  the host, the 640 independent frame pairs spread over a thread pool (OpenCV releases the GIL);
* the flow metrics (M-EPE / Fl-all / M-AE, :541-604) are evaluated where the flows already are,
  on the host, in the same fp32 torch expressions;
* steps whose policy cannot select anything (step <= 5, :412-417) skip the flow computation.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List

import numpy as np
import torch

from . import lib

_FARNEBACK = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def _flow_pair(args):
    import cv2
    a, b = args
    # the reference replicates the channel to RGB and converts back with COLOR_RGB2GRAY (:200-201),
    # which is the identity on uint8 for R=G=B (fixed-point weights sum to 1<<14; tests check all 256 values)
    return cv2.calcOpticalFlowFarneback(a, b, None, **_FARNEBACK)


def similarity_from_means(mean_epe: torch.Tensor, mean_outlier: torch.Tensor, mean_angle: torch.Tensor) -> float:
    """(:590-604) the three normalised error terms and their weighted sum, fp32 like the reference."""
    n_epe = torch.clamp(mean_epe / 10.0, 0.0, 1.0)
    n_fl = torch.clamp(mean_outlier / 0.5, 0.0, 1.0)
    n_ae = torch.clamp(mean_angle / 30.0, 0.0, 1.0)
    err = 0.45 * n_epe + 0.45 * n_fl + 0.1 * n_ae
    return torch.clamp(1.0 - err, 0.0, 1.0).item()


def flow_similarity(ref_flow: torch.Tensor, chan_flow: torch.Tensor) -> float:
    """[T-1, 2, H, W] fp32 flows -> similarity in [0,1] (:541-604)."""
    d = ref_flow - chan_flow
    epe = torch.sqrt((d ** 2).sum(dim=1) + 1e-8)
    dot = (ref_flow * chan_flow).sum(dim=1)
    rn = torch.sqrt((ref_flow ** 2).sum(dim=1) + 1e-8)
    cn = torch.sqrt((chan_flow ** 2).sum(dim=1) + 1e-8)
    cos = torch.clamp(dot / (rn * cn + 1e-8), -1.0, 1.0)
    ang = torch.acos(cos) * 180.0 / torch.pi
    outlier = (epe > 3.0) & (epe > rn * 0.05)
    return similarity_from_means(epe.mean(), outlier.float().mean(), ang.mean())


def selection_policy(scores, step: int) -> List[int]:
    """(:408-437) steps <= 5: none; 6..10: the single lowest; later: below mean - 0.625 std, clamped to [2, 6]."""
    s = np.array(scores)
    if step <= 10:
        out = np.argsort(s)[:(0 if step <= 5 else 1)].tolist()
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


class FlowChannelSelector:
    """``group``: when the sampler runs one process per GPU, the 16 channels are scored by different ranks (the host
    cores are shared by all ranks of the box) and the scores are summed into every rank - each rank then takes the
    same decision from the same 16 numbers."""

    def __init__(self, threads: int = 0, group=None, world: int = 1, rank: int = 0, device_flow=None):
        self.group, self.world, self.rank = group, world, rank
        # device_flow: Farneback + the flow metrics on the GPU (wf_farneback_u8 / wf_flow_metrics) for frames whose pyramid has
        # one or two levels (device_path_covers); other sizes, and WF_FLF_GPU=0, use OpenCV on host threads.  The kernels agree with OpenCV to ~3e-6 px; scores, selections and a 14-step guided trajectory are
        # identical through either path (tests/test_flow_gpu.py, tests/test_guided_loop_gpu.py).
        self.device_flow = (os.environ.get("WF_FLF_GPU", "1") == "1") if device_flow is None else bool(device_flow)
        self.threads = threads or max(1, min(32, (os.cpu_count() or 8) // max(world, 1)))
        self._pool = None
        self.last_scores = None

    @staticmethod
    def device_path_covers(H: int, W: int) -> bool:
        """wf_farneback_u8 restates the one-level pyramid (10 <= min side < 64: the 60 x 104 latent frames of 480p) and the
        two-level one with even sides (64 <= min side < 128: 90 x 160 at 720p, 88 x 160 in LongCat's refine pass)."""
        m = min(H, W)
        return 10 <= m < 64 or (64 <= m < 128 and H % 2 == 0 and W % 2 == 0)

    def _flows(self, u8: np.ndarray) -> torch.Tensor:
        """uint8 [C,T,H,W] -> fp32 [C, T-1, 2, H, W]."""
        C, T = u8.shape[:2]
        jobs = [(u8[c, t], u8[c, t + 1]) for c in range(C) for t in range(T - 1)]
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.threads)
        res = list(self._pool.map(_flow_pair, jobs, chunksize=4))
        fl = np.stack(res, axis=0).reshape(C, T - 1, *res[0].shape).transpose(0, 1, 4, 2, 3)
        t = torch.from_numpy(np.ascontiguousarray(fl)).float()
        if not torch.isfinite(t).all():
            ok = torch.isfinite(t)
            t = torch.where(ok, t, t[ok].mean() if ok.any() else torch.zeros(()))
        return t

    def scores(self, pred_x0: torch.Tensor, fused: torch.Tensor) -> List[float]:
        ref_u8 = lib.quantise_u8(fused.contiguous())
        pred_u8 = lib.quantise_u8(pred_x0.contiguous())
        nc = ref_u8.shape[1]
        c0, c1 = (self.rank * nc) // self.world, ((self.rank + 1) * nc) // self.world
        if self.device_flow and self.device_path_covers(*ref_u8.shape[-2:]) and c1 > c0:
            T, H, W = ref_u8.shape[2:]
            clips = torch.cat([ref_u8[0, c0:c1], pred_u8[0, c0:c1]]).contiguous()       # [2C', T, H, W] uint8, on the device
            flows = lib.farneback_u8(clips)                                              # [2C', T-1, H, W, 2]
            means = lib.flow_metrics(flows[:c1 - c0].contiguous(), flows[c1 - c0:].contiguous()).cpu()   # [C', 3]: one small D2H
            mine = [similarity_from_means(means[c, 0], means[c, 1], means[c, 2]) for c in range(c1 - c0)]
            return self._share(mine, nc, c0, c1, pred_x0.device)
        both = torch.stack([ref_u8[0, c0:c1], pred_u8[0, c0:c1]]).cpu().numpy()       # one D2H copy, [2,C',T,H,W]
        ref_fl, pred_fl = self._flows(both[0]), self._flows(both[1])
        mine = [flow_similarity(ref_fl[c], pred_fl[c]) for c in range(ref_fl.shape[0])]
        return self._share(mine, nc, c0, c1, pred_x0.device)

    def _share(self, mine, nc: int, c0: int, c1: int, device) -> List[float]:
        if self.world > 1:
            import torch.distributed as dist
            t = torch.zeros(nc, dtype=torch.float64)
            t[c0:c1] = torch.tensor(mine, dtype=torch.float64)
            t = t.to(device)
            dist.all_reduce(t, group=self.group)                 # x + 0 is exact: every rank holds the same 16 scores
            mine = t.cpu().tolist()
        self.last_scores = mine
        return self.last_scores

    def select(self, pred_x0: torch.Tensor, fused: torch.Tensor, step: int) -> List[int]:
        if step <= 5:
            return []
        return selection_policy(self.scores(pred_x0, fused), step)
