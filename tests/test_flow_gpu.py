"""Device Farneback flow + flow metrics (wf_farneback_u8 / wf_flow_metrics) against the oracle (numpy restatement pinned to
OpenCV) and against OpenCV itself, and the FLF channel selector's scores through either path."""
import time

import numpy as np
import pytest
import torch

from oracle import farneback as fb

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
ARGS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def _smooth(seed, shape=(60, 104), sigma=4.0):
    r = np.random.default_rng(seed)
    a = cv2.GaussianBlur((r.random(shape) * 255).astype(np.float32), (0, 0), sigma)
    return (a - a.min()) / (a.max() - a.min()) * 255


def _clips(n, T, shape=(60, 104)):
    rng = np.random.default_rng(99)
    u8 = np.zeros((n, T) + shape, np.uint8)
    for c in range(n):
        base = _smooth(100 + c, shape)
        for t in range(T):
            if c % 3 == 0:
                u8[c, t] = np.roll(base, 2 * t, axis=1).astype(np.uint8)
            elif c % 3 == 1:
                u8[c, t] = np.roll(np.roll(base, t, axis=0), -t, axis=1).astype(np.uint8)
            else:
                u8[c, t] = (rng.random(shape) * 255).astype(np.uint8)
    return u8


@pytest.mark.parametrize("shape", [(60, 104), (40, 56), (90, 160), (88, 160), (64, 72)])     # the last three: two-level pyramids
def test_device_flow_matches_oracle_and_opencv(cuda, shape):
    from worldforge_b200 import lib
    u8 = _clips(6, 4, shape)
    got = lib.farneback_u8(torch.from_numpy(u8).to(cuda)).cpu().numpy()
    assert got.shape == (6, 3) + shape + (2,)
    for c in range(6):
        for t in range(3):
            want = fb.farneback(u8[c, t], u8[c, t + 1])
            ref = cv2.calcOpticalFlowFarneback(u8[c, t], u8[c, t + 1], None, **ARGS)
            np.testing.assert_allclose(got[c, t], want, rtol=0, atol=1e-4)      # FMA contraction in the double sums only
            # two levels: OpenCV's filter engine rounds the sigma-0.5 blur differently (oracle/farneback.py: ~1e-4 px)
            np.testing.assert_allclose(got[c, t], ref, rtol=0, atol=1e-4 if min(shape) < 64 else 1e-3)


def test_device_metrics_match_the_host_expressions(cuda):
    from worldforge_b200 import flf_select, lib
    g = torch.Generator().manual_seed(3)
    a = torch.randn(5, 7, 20, 24, 2, generator=g) * 3
    b = a + torch.randn(5, 7, 20, 24, 2, generator=g) * torch.tensor([0.1, 1.0, 3.0, 6.0, 0.0]).view(5, 1, 1, 1, 1)
    m = lib.flow_metrics(a.to(cuda), b.to(cuda)).cpu()
    for c in range(5):
        ra, rb = a[c].permute(0, 3, 1, 2), b[c].permute(0, 3, 1, 2)           # [T-1, 2, H, W] as the host path holds them
        want = flf_select.flow_similarity(ra, rb)
        got = flf_select.similarity_from_means(m[c, 0], m[c, 1], m[c, 2])
        assert abs(got - want) < 2e-6, (c, got, want)


@pytest.mark.parametrize("shape", [(60, 104), (90, 160)])
def test_selector_scores_device_vs_opencv(cuda, shape):
    """The whole scoring call on 16 channels x 21 latent frames (60 x 104 at 480p: one pyramid level; 90 x 160 at 720p: two):
    device path vs OpenCV path, and its timing."""
    from worldforge_b200 import flf_select
    u8 = _clips(32, 21, shape)
    # latents whose min-max quantisation reproduces the clips: x = u8 / 255 spans [0, 1] in both tensors
    ref = torch.from_numpy(u8[:16].astype(np.float32) / 255.0).unsqueeze(0).to(cuda)
    pred = torch.from_numpy(u8[16:].astype(np.float32) / 255.0).unsqueeze(0).to(cuda)
    ref[0, 0, 0, 0, 0], ref[0, 0, 0, 0, 1] = 0.0, 1.0
    pred[0, 0, 0, 0, 0], pred[0, 0, 0, 0, 1] = 0.0, 1.0
    host = flf_select.FlowChannelSelector(device_flow=False)
    dev = flf_select.FlowChannelSelector(device_flow=True)
    s_host = host.scores(pred, ref)
    s_dev = dev.scores(pred, ref)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); dev.scores(pred, ref); torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); host.scores(pred, ref); t_host = time.perf_counter() - t0
    print(f"\nFLF scoring {shape}: device {t_dev * 1e3:.1f} ms, OpenCV on {host.threads} host threads {t_host * 1e3:.1f} ms")
    assert dev.device_path_covers(*shape)
    np.testing.assert_allclose(s_dev, s_host, rtol=0, atol=2e-5)
    for step in (7, 12, 30):
        assert flf_select.selection_policy(s_dev, step) == flf_select.selection_policy(s_host, step)


def test_longcat_selector_device_vs_opencv(cuda):
    """LongCat's variant of the scoring (per-channel min-max, outliers by OR, weights 0.4 / 0.4 / 0.2;
    scheduling_flow_match_euler_discrete.py:205-244, 286-379): device path vs OpenCV on host threads, same scores to 2e-5 and
    the same channel choices at every policy stage."""
    from worldforge_b200 import longcat_pipeline as wlp, lib
    u8 = _clips(32, 13, (60, 104))
    enc = torch.from_numpy(u8[:16].astype(np.float32) / 255.0).unsqueeze(0).to(cuda) * 3.0 - 1.0
    pred = torch.from_numpy(u8[16:].astype(np.float32) / 255.0).unsqueeze(0).to(cuda) * 2.0 + 0.5
    host, dev = wlp.LongCatChannelSelector(device_flow=False), wlp.LongCatChannelSelector(device_flow=True)
    for step, distill in ((2, False), (8, False), (5, True)):
        a = host.select(pred, enc, step, distill, 3)
        b = dev.select(pred, enc, step, distill, 3)
        np.testing.assert_allclose(dev.last_scores, host.last_scores, rtol=0, atol=2e-5)
        assert a == b and len(a) >= 1
    g = torch.Generator().manual_seed(3)
    fa = torch.randn(3, 7, 20, 24, 2, generator=g) * 3
    fb_ = fa + torch.randn(3, 7, 20, 24, 2, generator=g) * torch.tensor([0.1, 2.0, 6.0]).view(3, 1, 1, 1, 1)
    m = lib.flow_metrics(fa.to(cuda), fb_.to(cuda), outlier_or=True).cpu()
    for c in range(3):
        want = wlp.flow_similarity(fa[c].permute(0, 3, 1, 2), fb_[c].permute(0, 3, 1, 2))
        assert abs(wlp.similarity_from_means(float(m[c, 0]), float(m[c, 1]), float(m[c, 2])) - want) < 2e-6
