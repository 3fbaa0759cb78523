"""The numpy restatement of OpenCV's Farneback flow (oracle/farneback.py) pinned against OpenCV itself, on the frame size
and parameters the FLF channel selector uses (reference scheduling_unipc_multistep_clean.py:220-224), and the selector's
decisions taken from either flow source compared."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from oracle import farneback as fb
from worldforge_b200 import flf_select

ARGS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def _smooth(seed, shape=(60, 104), sigma=4.0):
    r = np.random.default_rng(seed)
    a = cv2.GaussianBlur((r.random(shape) * 255).astype(np.float32), (0, 0), sigma)
    return (a - a.min()) / (a.max() - a.min()) * 255


CASES = {
    "shift2": lambda: (_smooth(0).astype(np.uint8), np.roll(_smooth(0), 2, axis=1).astype(np.uint8)),
    "shift_diag": lambda: (_smooth(1).astype(np.uint8), np.roll(np.roll(_smooth(1), 1, axis=0), -3, axis=1).astype(np.uint8)),
    "noise": lambda: ((np.random.default_rng(2).random((60, 104)) * 255).astype(np.uint8),
                      (np.random.default_rng(3).random((60, 104)) * 255).astype(np.uint8)),
    "constant": lambda: (np.full((60, 104), 37, np.uint8), np.full((60, 104), 37, np.uint8)),
    "unrelated_smooth": lambda: (_smooth(4, sigma=2.0).astype(np.uint8), _smooth(5, sigma=2.0).astype(np.uint8)),
    "other_size": lambda: (_smooth(6, (40, 56)).astype(np.uint8), np.roll(_smooth(6, (40, 56)), 1, axis=0).astype(np.uint8)),
    # 720p latent frames: two pyramid levels (45 x 80, then 90 x 160)
    "720p_shift": lambda: (_smooth(7, (90, 160)).astype(np.uint8), np.roll(np.roll(_smooth(7, (90, 160)), 2, axis=0), -3, axis=1).astype(np.uint8)),
    "720p_noise": lambda: ((np.random.default_rng(8).random((90, 160)) * 255).astype(np.uint8),
                           (np.random.default_rng(9).random((90, 160)) * 255).astype(np.uint8)),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_restatement_matches_opencv(case):
    a, b = CASES[case]()
    ref = cv2.calcOpticalFlowFarneback(a, b, None, **ARGS)
    got = fb.farneback(a, b)
    assert got.shape == ref.shape and got.dtype == np.float32
    # double-precision running sums in OpenCV vs direct sums here: agreement to a few float32 ulps of the flow; the
    # two-level case also carries OpenCV's filter-engine rounding of the sigma-0.5 blur (1 ulp of a 0..255 pixel)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-4 if case.startswith("720p") else 2e-5)


def test_pieces_against_a_one_iteration_call():
    """iterations=1 isolates polynomial expansion + the first matrix build + one box solve."""
    a, b = CASES["shift2"]()
    ref = cv2.calcOpticalFlowFarneback(a, b, None, **dict(ARGS, iterations=1))
    got = fb.farneback(a, b, iterations=1)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)


def test_deeper_pyramids_are_refused():
    with pytest.raises(NotImplementedError):
        fb.farneback(np.zeros((180, 320), np.uint8), np.zeros((180, 320), np.uint8))       # three levels
    with pytest.raises(NotImplementedError):
        fb.farneback(np.zeros((91, 160), np.uint8), np.zeros((91, 160), np.uint8))         # two levels, odd side


def test_channel_selection_is_the_same_from_either_flow():
    """16 'channels' of 6 frames: the similarity scores and the selected channels computed from the restated flows equal
    those computed from OpenCV's (what a device implementation will be gated on)."""
    rng = np.random.default_rng(7)
    T, C = 6, 16
    ref_clip = np.stack([np.stack([np.roll(_smooth(10 + c), 2 * t, axis=1) for t in range(T)]) for c in range(C)]).astype(np.uint8)
    # candidates: some channels move like the reference, some differently, some are noise
    cand = ref_clip.copy()
    for c in range(C):
        if c % 3 == 1:
            cand[c] = np.stack([np.roll(_smooth(10 + c), -t, axis=0) for t in range(T)]).astype(np.uint8)
        elif c % 3 == 2:
            cand[c] = (rng.random((T, 60, 104)) * 255).astype(np.uint8)

    def flows(clip, fn):
        return torch.from_numpy(np.stack([np.stack([fn(clip[c, t], clip[c, t + 1]) for t in range(T - 1)]) for c in range(C)])
                                .transpose(0, 1, 4, 2, 3).copy())

    cvf = lambda a, b: cv2.calcOpticalFlowFarneback(a, b, None, **ARGS)
    s_cv = [flf_select.flow_similarity(r, c) for r, c in zip(flows(ref_clip, cvf), flows(cand, cvf))]
    s_or = [flf_select.flow_similarity(r, c) for r, c in zip(flows(ref_clip, fb.farneback), flows(cand, fb.farneback))]
    np.testing.assert_allclose(s_or, s_cv, rtol=0, atol=1e-5)
    for step in (7, 12, 30):
        assert flf_select.selection_policy(s_or, step) == flf_select.selection_policy(s_cv, step)
