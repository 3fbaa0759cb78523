"""FLF channel selection: which latent channels keep the model's own prediction.

Host side of the reference's ``VideoMotionPCASelector`` (reference
utils/scheduling_unipc_multistep_clean.py:338-607).  The reference scores each of the 16
latent channels by comparing dense optical flow (OpenCV Farneback, :220-224) of the fused
latents with that of the model's prediction, after quantising each channel to uint8.  Here

* the min-max quantisation of both 16-channel tensors runs on the GPU (``wf_quantise_u8``,
  two passes over 2 MB) so 4 MB of uint8 cross PCIe instead of 16 MB of fp32 in 32 copies;
* Farneback itself stays OpenCV on the host - it is the reference's own third-party dependency
  for this step and its discrete outcome (an argsort over 16 scores) must not drift
  (SURVEY.md §7 "hard parts", §8f item 1) - but the 640 independent frame pairs are spread over
  a thread pool (OpenCV releases the GIL);
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
    n_epe = torch.clamp(epe.mean() / 10.0, 0.0, 1.0)
    n_fl = torch.clamp(outlier.float().mean() / 0.5, 0.0, 1.0)
    n_ae = torch.clamp(ang.mean() / 30.0, 0.0, 1.0)
    err = 0.45 * n_epe + 0.45 * n_fl + 0.1 * n_ae
    return torch.clamp(1.0 - err, 0.0, 1.0).item()


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

    def __init__(self, threads: int = 0, group=None, world: int = 1, rank: int = 0):
        self.group, self.world, self.rank = group, world, rank
        self.threads = threads or max(1, min(32, (os.cpu_count() or 8) // max(world, 1)))
        self._pool = None
        self.last_scores = None

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
        both = torch.stack([ref_u8[0, c0:c1], pred_u8[0, c0:c1]]).cpu().numpy()       # one D2H copy, [2,C',T,H,W]
        ref_fl, pred_fl = self._flows(both[0]), self._flows(both[1])
        mine = [flow_similarity(ref_fl[c], pred_fl[c]) for c in range(ref_fl.shape[0])]
        if self.world > 1:
            import torch.distributed as dist
            t = torch.zeros(nc, dtype=torch.float64)
            t[c0:c1] = torch.tensor(mine, dtype=torch.float64)
            t = t.to(pred_x0.device)
            dist.all_reduce(t, group=self.group)                 # x + 0 is exact: every rank holds the same 16 scores
            mine = t.cpu().tolist()
        self.last_scores = mine
        return self.last_scores

    def select(self, pred_x0: torch.Tensor, fused: torch.Tensor, step: int) -> List[int]:
        if step <= 5:
            return []
        return selection_policy(self.scores(pred_x0, fused), step)
