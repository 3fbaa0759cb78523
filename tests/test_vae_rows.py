"""Host logic of the VAE's row sharding (no GPU): the backwards walk that finds which input rows a rank's output rows
depend on, checked against torch convolutions standing in for the layers, and the gather of uneven row slabs over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from worldforge_b200 import vae as wvae


def _layer(kind, x, w):
    """A stand-in with the same spatial footprint as the VAE layer kinds (x: [1,1,H,W])."""
    conv = lambda t: F.conv2d(t, w, padding=1)
    if kind in ("conv", "head"):
        return conv(x)
    if kind == "res":
        return x + conv(torch.tanh(conv(x)))
    if kind in ("up2d", "up3d"):
        return conv(F.interpolate(x, scale_factor=2.0, mode="nearest-exact"))
    if kind in ("down2d", "down3d"):
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, stride=2)
    raise ValueError(kind)


def _run(seg, x, w):
    for kind, *_ in seg:
        x = _layer(kind, x, w)
    return x


ENC = [("conv", "c", 0, 0), ("res", "r", 0, 0), ("res", "r", 0, 0), ("down2d", "d", 0, 0), ("res", "r", 0, 0), ("down3d", "d", 0, 0),
       ("res", "r", 0, 0), ("down3d", "d", 0, 0)]
DEC = [("up3d", "u", 0, 0), ("res", "r", 0, 0), ("res", "r", 0, 0), ("up3d", "u", 0, 0), ("res", "r", 0, 0), ("up2d", "u", 0, 0),
       ("res", "r", 0, 0), ("head", "h", 0, 0)]


class _StandIn(wvae.WfWanVAE):
    """WfWanVAE's row-slab driver (_run_rows) over torch stand-ins of the layers; activations are [1, H, W, 1]."""

    def __init__(self, w):
        self.wt = w

    def _run(self, plan, x, final_planar=None):
        y = x.permute(0, 3, 1, 2)
        for kind, *_ in plan:
            y = _layer(kind, y, self.wt)
        return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("seg,h_in,out_scale", [(ENC, 96, 1), (DEC, 12, 8)])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_row_slabs_reproduce_the_full_evaluation(seg, h_in, out_scale, world):
    """A rank's slab - the rows need[0] with zero padding at the slab edges (what the conv kernels do), re-cut before
    every resampling layer - yields the rank's output rows exactly."""
    torch.manual_seed(0)
    x = torch.randn(1, h_in, 16, 1, dtype=torch.float64)
    m = _StandIn(torch.randn(1, 1, 3, 3, dtype=torch.float64))
    full = m._run(seg, x)
    b = m.row_bounds(12, world)
    for r in range(world):
        lo, hi = out_scale * b[r], out_scale * b[r + 1]
        need, hs = m._needed_rows(seg, (lo, hi), h_in)
        assert hs[-1] == full.shape[1] and need[-1] == (lo, hi)
        a, e = need[0]
        y, first = m._run_rows(seg, x[:, a:e], a, need)
        assert first == lo and torch.equal(y, full[:, lo:hi]), (r, need)
        for kind, (nlo, nhi) in zip([s[0] for s in seg], need):
            if kind in m.DOWNS:
                assert nlo % 2 == 0 and (nhi - nlo) % 2 == 0      # space-to-depth needs aligned, even slabs


def test_row_bounds_partition():
    for h in (60, 90, 7, 8):
        for world in (1, 2, 4, 8):
            if world > h:
                continue
            b = wvae.WfWanVAE.row_bounds(h, world)
            assert b[0] == 0 and b[-1] == h and len(b) == world + 1
            sizes = [b[i + 1] - b[i] for i in range(world)]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


def _store_file():
    import tempfile
    fd, path = tempfile.mkstemp(prefix="wf_gloo_"); os.close(fd); os.unlink(path)
    return path


def _worker(rank, world, store, ret):
    dist.init_process_group("gloo", init_method=f"file://{store}", rank=rank, world_size=world)   # no TCP port to race for
    try:
        m = wvae.WfWanVAE.__new__(wvae.WfWanVAE)
        m.enable_row_sharding()
        assert (m.shard.world, m.shard.rank) == (world, rank)
        full = torch.arange(2 * 7 * 3, dtype=torch.float32).reshape(2, 7, 3)
        b = m.row_bounds(7, world)                                   # uneven: 3 + 4 rows
        got = m._all_gather_rows(full[:, b[rank]:b[rank + 1]].contiguous(), 1, b)
        ret[rank] = torch.equal(got, full)
    finally:
        dist.destroy_process_group()


def test_gather_of_uneven_row_slabs_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _store_file(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


class _StagedStandIn(_StandIn):
    """Stand-ins for the pieces sharded_stages drives besides the row slabs: a full-frame 'attention' (every output pixel
    depends on the whole frame - it is only right if the rows were gathered before and the frames split after) and the
    final 1x1 convolution."""
    z_dim, spatial_scale = 1, 8

    def __init__(self, w):
        super().__init__(w)
        self.enc_plan = ENC + [("res", "r", 0, 0), ("attn", "a", 0, 0), ("res", "r", 0, 0), ("head", "h", 0, 0)]
        self.dec_plan = []

    def _attn(self, x, name, c):
        return x + x.mean(dim=(1, 2), keepdim=True) * torch.arange(1, x.shape[1] + 1, dtype=x.dtype).view(1, -1, 1, 1)

    def _conv(self, x, wname, taps, cout, **kw):
        return 2.0 * x


@pytest.mark.parametrize("cuts", [False, True])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_encode_stages_rows_frames_rows(world, cuts, monkeypatch):
    """sharded_stages('enc'): [rows | frames | rows] with a gather between stages reproduces the unsharded evaluation;
    with ``level_cuts`` the row stages additionally end after every downsampling layer (one stage per resolution level)."""
    monkeypatch.setattr(wvae.lib, "planar_to_cl", lambda src, Cp, round_tf32=False: src.permute(1, 2, 3, 0).contiguous())
    torch.manual_seed(1)
    m = _StagedStandIn(torch.randn(1, 1, 3, 3) * 0.3)               # fp32: the first stage casts the video like the engine does
    video = torch.randn(1, 5, 96, 16)                               # planar [C=1, F, H, W]
    # the reference evaluation, layer by layer (attention in its place)
    x = video.permute(1, 2, 3, 0)
    for layer in m.enc_plan:
        x = m._attn(x, "a", 0) if layer[0] == "attn" else m._run([layer], x)
    want = m._conv(x, "conv1", None, 2)
    m.level_cuts = cuts
    stages = m.sharded_stages("enc", video, world)
    n_down = sum(1 for l in m.enc_plan if l[0] in m.DOWNS)
    assert len(stages) == 3 + (n_down if cuts else 0) and n_down >= 2
    full = video
    for stage in stages:
        parts, dim = [], None
        for r in range(world):
            part, dim, bounds = stage(r, full)
            assert part.shape[dim] == bounds[r + 1] - bounds[r]
            parts.append(part)
        full = wvae.assemble_rows(parts, dim)
    torch.testing.assert_close(full, want, rtol=1e-5, atol=1e-5)    # torch's CPU convolution may block a cropped input differently
