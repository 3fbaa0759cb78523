"""Input preparation: the oracle against the reference's own soften_mask (committed fixture + live),
the device kernels against the oracle (bit for bit)."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin
from oracle import make_inputs_golden as mig

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "inputs_golden.pt"), weights_only=False)
CASES = ((15, "sine"), (7, "linear"), (10, "exponential"), (4, "cosine"))


def test_oracle_soften_mask_matches_reference_fixture():
    arr = oin.stack_masks(GOLD["masks_u8"].numpy())
    assert np.array_equal(GOLD["masks_u8"].numpy(), mig.masks())
    for td, kind in CASES:
        got = oin.soften_mask(arr, td, kind)
        assert got.dtype == np.float32
        assert np.array_equal(got, GOLD[f"soft_{td}_{kind}"].numpy()), (td, kind)
    soft = GOLD["soft_15_sine"].numpy()
    assert 0.0 < soft[0][soft[0] > 0].min() < 0.2 and np.array_equal(soft[3], np.ones_like(soft[3])) and not soft[4].any()


@pytest.mark.skipif(not os.path.exists(mig.REF), reason="/root/reference is not mounted")
def test_oracle_soften_mask_matches_live_reference():
    ref = oin.reference_soften_mask(mig.REF)
    rng = np.random.default_rng(5)
    arr = (rng.random((3, 50, 64)) > 0.3).astype(np.float64)
    arr[1, 20:30, 20:40] = 1.0
    for td, kind in CASES + ((2.5, "sine"),):
        assert np.array_equal(oin.soften_mask(arr, td, kind), ref(arr, td, kind))


def test_ramp_table_is_the_reference_expression():
    from worldforge_b200 import inputs
    for td, kind in CASES + ((2.5, "sine"),):
        table, radius, max_d2 = inputs._ramp_table(td, kind)
        assert radius == int(np.floor(td)) and max_d2 == int(np.floor(td * td)) and len(table) == max_d2 + 1
        d = np.sqrt(np.arange(max_d2 + 1, dtype=np.float64))
        assert np.array_equal(table, oin.smooth_transition(d / td, kind).astype(np.float32))
    assert np.array_equal(inputs._U8_TO_UNIT.numpy(), (np.arange(256) / 255.0).astype(np.float32))


@pytest.mark.gpu
def test_device_soften_mask_is_bit_identical(cuda):
    from worldforge_b200 import inputs
    mu8 = GOLD["masks_u8"].to(cuda)
    for td, kind in CASES:
        got = inputs.soften_mask(mu8, td, kind)
        assert torch.equal(got.cpu(), GOLD[f"soft_{td}_{kind}"]), (td, kind)
    # a full-size clip with ragged edges against the oracle (not a multiple of the 32-pixel tile), fractional distance
    rng = np.random.default_rng(9)
    arr = (rng.random((4, 123, 211)) > 0.004).astype(np.float32)
    arr[2, 40:80, 100:160] = 0.0
    for td, kind in ((15, "sine"), (2.5, "cosine"), (31, "linear")):
        want = oin.soften_mask(arr.astype(np.float64), td, kind)
        got = inputs.soften_mask(torch.from_numpy(arr).to(cuda), td, kind)
        assert torch.equal(got.cpu(), torch.from_numpy(want)), (td, kind)
    assert inputs.prepare_mask(mu8).shape == (1, 1) + tuple(mu8.shape)


@pytest.mark.gpu
def test_device_frame_stacking(cuda):
    from worldforge_b200 import inputs
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (5, 37, 53, 3), generator=g, dtype=torch.uint8)
    got = inputs.clip_from_frames(frames.to(cuda))
    assert torch.equal(got.cpu(), torch.from_numpy(oin.stack_frames(frames.numpy())))
    want = torch.stack([torch.tensor(np.array(f)).permute(2, 0, 1).float() / 255.0 for f in frames.numpy()]).unsqueeze(0).permute(0, 2, 1, 3, 4)
    assert torch.equal(got.cpu(), want)                              # the reference's own expression (:232-238)


# ---- fuse_latents' resize branch (scheduling_unipc_multistep_clean.py:1297-1371) ----------------------------------------------
PRESIZE = torch.load(os.path.join(os.path.dirname(__file__), "golden", "presize_golden.pt"), weights_only=False)


def _oracle_fused(clip, mask):
    from oracle import unipc
    dec, x0, _ = mig.presize_inputs()
    vae = mig.RecordingVAE(dec)
    unipc.OracleUniPC(flow_shift=3.0).fuse_latents(x0, clip, mask, vae=vae)
    return vae.fused


def test_oracle_resize_branch_matches_reference_fixture():
    """What the reference's own fuse_latents hands to vae.encode when the warped clip / mask are not the decoded clip's size
    (bilinear clip, nearest mask, first mask channel), bit for bit."""
    _, _, cases = mig.presize_inputs()
    assert set(cases) == set(PRESIZE)
    for name, (clip, mask) in cases.items():
        assert torch.equal(_oracle_fused(clip, mask), PRESIZE[name]), name


@pytest.mark.skipif(not os.path.exists(mig.REF), reason="/root/reference is not mounted")
def test_oracle_resize_branch_matches_live_reference_and_fails_where_it_fails():
    live = mig.ref_presize()
    for name, want in PRESIZE.items():
        assert torch.equal(live[name], want), name
    from oracle import ref_shim
    sm = ref_shim.load_scheduler_module()
    s = sm.UniPCMultistepScheduler(num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction",
                                   use_flow_sigmas=True, flow_shift=3.0)
    dec, x0, _ = mig.presize_inputs()
    bad = {"frames": (torch.rand(1, 3, 3, 32, 48), torch.rand(1, 1, 3, 32, 48)),          # temporal branch: F.interpolate rejects it
           "mask_only": (torch.rand(1, 3, 5, 32, 48), torch.rand(1, 1, 5, 16, 24))}       # F unbound outside the clip branch
    for name, (clip, mask) in bad.items():
        with pytest.raises((ValueError, UnboundLocalError)):
            s.fuse_latents(x0, clip, mask, vae=mig.RecordingVAE(dec))
        with pytest.raises(ValueError):
            _oracle_fused(clip, mask)


def test_engine_presize_equals_oracle_and_scheduler_keeps_the_resized_pair(monkeypatch):
    """worldforge_b200.inputs.presize_guidance against the oracle; WfUniPCScheduler.fuse_latents' host path (kernels stood in
    for by their torch expressions) reproduces the reference fixture and resizes a given pair once."""
    import contextlib
    from oracle import flf
    from worldforge_b200 import inputs, lib, scheduler as wsched
    dec, x0, cases = mig.presize_inputs()
    for name, (clip, mask) in cases.items():
        a, b = inputs.presize_guidance(clip, mask, dec.shape)
        c, d = flf.presize_guidance(clip, mask, dec.shape)
        assert torch.equal(a, c) and torch.equal(b, d) and a.shape == dec.shape and b.shape == (1, 1) + dec.shape[2:], name
    for clip, mask in ((torch.rand(1, 3, 3, 32, 48), torch.rand(1, 1, 3, 32, 48)), (torch.rand(1, 3, 5, 32, 48), torch.rand(1, 1, 5, 16, 24))):
        with pytest.raises(ValueError):
            inputs.presize_guidance(clip, mask, dec.shape)

    calls = []
    real = inputs.presize_guidance
    monkeypatch.setattr(inputs, "presize_guidance", lambda *a: (calls.append(1), real(*a))[1])
    monkeypatch.setattr(lib, "latent_denorm", lambda x, m, s: x.float())
    monkeypatch.setattr(lib, "flf_blend", lambda d, r, m: (2.0 * r - 1.0) * m + d * (1 - m))
    monkeypatch.setattr(lib, "latent_norm_replace", lambda enc, x, m, s, ch: enc)
    monkeypatch.setattr(lib, "phase", lambda name: contextlib.nullcontext())
    s = wsched.WfUniPCScheduler(flow_shift=3.0)
    s.set_timesteps(4)
    for name, (clip, mask) in cases.items():
        vae = mig.RecordingVAE(dec)
        n = len(calls)
        for _ in range(3):
            s.fuse_latents(x0, clip, mask, vae=vae)
            assert torch.equal(vae.fused, PRESIZE[name]), name
        assert len(calls) == n + 1, name                       # resized on the first call only
        mask.add_(0.0)                                          # an in-place edit of a source invalidates the kept pair
        s.fuse_latents(x0, clip, mask, vae=vae)
        assert len(calls) == n + 2


# ---- the on-disk contract: a warp folder of frames and mask_* images (infer_worldforge.py:65-102, :208-251) --------------------
def _write_case(tmp_path, n_frames=5, n_masks=3, size=(50, 36)):
    from PIL import Image
    rng = np.random.default_rng(11)
    for i in range(n_frames):
        ext = "jpg" if i == 1 else "png"                     # mixed extensions sort by the whole path, as in the reference
        Image.fromarray(rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)).save(tmp_path / f"warp_{i:03d}.{ext}")
    for i in range(n_masks):
        m = np.zeros((size[1], size[0]), np.uint8)
        m[5 + i:25, 8:30 + i] = 255
        Image.fromarray(m).save(tmp_path / f"mask_{i:03d}.png")
    (tmp_path / "notes.txt").write_text("not an image")
    return str(tmp_path)


def test_case_folder_listing_and_mask_padding(tmp_path):
    from worldforge_b200 import inputs
    d = _write_case(tmp_path)
    frames, masks = inputs.list_case_folder(d)
    assert [os.path.basename(f) for f in frames] == ["warp_000.png", "warp_001.jpg", "warp_002.png", "warp_003.png", "warp_004.png"]
    assert [os.path.basename(f) for f in masks] == ["mask_000.png", "mask_001.png", "mask_002.png"]
    fr, mk, first = inputs.read_case_folder(d)
    assert len(fr) == len(mk) == 5 and first is fr[0] and fr[0].mode == "RGB" and mk[0].mode == "L"
    assert mk[3] is mk[2] and mk[4] is mk[2]                 # fewer masks than frames: the last one repeated (:95-97)
    with pytest.raises(ValueError):
        inputs.list_case_folder(str(tmp_path / "missing"))
    assert inputs.target_size(1280, 720, 480 * 832) == (832, 464) and inputs.target_size(960, 512, 480 * 832) == (864, 448)
    assert inputs.target_size(1280, 720, 720 * 1280) == (1280, 720) and inputs.target_size(832, 480) == (832, 480)


@pytest.mark.skipif(not os.path.exists(mig.REF), reason="/root/reference is not mounted")
def test_case_folder_matches_the_reference_entry_script(tmp_path):
    """read_frames_from_directory and the size statements of infer_worldforge.py, taken from its source, against the engine's
    loader - on a synthetic folder (mixed extensions, fewer masks than frames, no masks, more masks) and on the reference's
    own truck case; then the whole of :225-251 (PIL resize, stacking, softening) against load_case's composition."""
    from worldforge_b200 import inputs
    ref_read = oin.reference_read_frames(mig.REF)
    same = lambda a, b: len(a) == len(b) and all(x.size == y.size and x.mode == y.mode and x.tobytes() == y.tobytes() for x, y in zip(a, b))
    for k, (nf, nm) in enumerate(((5, 3), (4, 0), (3, 6))):
        sub = tmp_path / f"case{k}"
        sub.mkdir()
        d = _write_case(sub, nf, nm)
        (f0, m0, first0), (f1, m1, first1) = ref_read(d), inputs.read_case_folder(d)
        assert same(f0, f1) and same(m0, m1) and same([first0], [first1]) and len(m1) == nf
    truck = os.path.join(os.path.dirname(os.path.dirname(mig.REF)), "test_case", "truck", "imgs")
    fr, mk = inputs.list_case_folder(truck)
    assert len(fr) == len(mk) == 49
    rng = np.random.default_rng(3)
    for w, h in [(1280, 720), (960, 512), (720, 1280), (1000, 1000), (641, 479)] + [tuple(int(v) for v in rng.integers(200, 2000, 2)) for _ in range(20)]:
        for area in (480 * 832, 720 * 1280):
            assert inputs.target_size(w, h, area) == oin.reference_target_size(mig.REF, w, h, area), (w, h, area)

    d = _write_case(tmp_path / "case0", 5, 3, size=(100, 60))
    got = inputs.load_case(d, max_area=64 * 96, transition_distance=6, decay_type="sine",
                           _to_clip=lambda u8: torch.from_numpy(oin.stack_frames(u8.numpy())),
                           _to_mask=lambda u8: torch.from_numpy(oin.soften_mask(oin.stack_masks(u8.numpy()), 6, "sine"))[None, None])
    frames, masks, first = ref_read(d)
    width, height = oin.reference_target_size(mig.REF, first.width, first.height, 64 * 96)
    assert (got["width"], got["height"], got["num_frames"]) == (width, height, 5) and got["image"].size == (width, height)
    resized = [f.resize((width, height)) for f in frames]                                                      # :231
    video = torch.stack([torch.tensor(np.array(f)).permute(2, 0, 1).float() / 255.0 for f in resized])         # :232-235
    assert torch.equal(got["video_ref"], video.unsqueeze(0).permute(0, 2, 1, 3, 4))                            # :236
    marr = np.stack([np.array(m.resize((width, height))) / 255.0 for m in masks])                              # :245-246
    marr = oin.reference_soften_mask(mig.REF)(marr, 6, "sine")                                                 # :248-249
    assert torch.equal(got["mask"], torch.from_numpy(marr).unsqueeze(0).unsqueeze(0))                          # :251
    assert got["image"].tobytes() == first.resize((width, height)).tobytes()                                   # :222
