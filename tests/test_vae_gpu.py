"""Parity of the CUDA 3D-VAE (implicit-GEMM convs on tcgen05/tf32 + fused norm kernels) with the oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_vae

pytestmark = pytest.mark.gpu


def g(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("cin,cout,T,H,W", [(32, 96, 3, 16, 24), (96, 96, 2, 20, 16), (192, 384, 2, 8, 8), (16, 384, 3, 8, 16),
                                            (4, 96, 3, 16, 16), (96, 3, 2, 16, 16)])
def test_causal_conv3d(cuda, cin, cout, T, H, W):
    from worldforge_b200 import lib, vae as wvae
    x = torch.randn(cin, T, H, W, generator=g(1))
    w = torch.randn(cout, cin, 3, 3, 3, generator=g(2)) / (cin * 27) ** 0.5
    b = torch.randn(cout, generator=g(3)) * 0.1
    want = wan_vae.causal_conv3d(x, w, b)                       # [cout, T, H, W]
    xc = lib.planar_to_cl(x.to(cuda), (cin + 3) // 4 * 4)
    out = torch.empty(T, H, W, cout, device=cuda)
    lib.conv_tf32(xc, wvae._w_conv3d(w).to(cuda), b.to(cuda), wvae.TAPS_333, out, T=T, H=H, W=W, Cout=cout,
                  tile_w=16 if W % 16 == 0 else 8)
    got = out.permute(3, 0, 1, 2).cpu()
    assert rel(got, want) < 2e-3, rel(got, want)      # tf32 operands (10-bit mantissa), fp32 accumulation
    # residual add fused in the epilogue
    r = torch.randn(T, H, W, cout, generator=g(4)).to(cuda)
    out2 = torch.empty_like(out)
    lib.conv_tf32(xc, wvae._w_conv3d(w).to(cuda), b.to(cuda), wvae.TAPS_333, out2, T=T, H=H, W=W, Cout=cout, resid=r,
                  tile_w=16 if W % 16 == 0 else 8)
    torch.testing.assert_close(out2, out + r, rtol=1e-6, atol=1e-6)


def test_upsample_downsample_blocks(cuda):
    from worldforge_b200 import vae as wvae
    cfg = wan_vae.VaeConfig(dim=8)
    P = wan_vae.init_params(cfg, 5)
    m = wvae.WfWanVAE(P, cuda, dim=8)
    # decoder.upsamples.3 is the first up3d (32 -> 16 channels), encoder.downsamples.5 the first down3d (16 ch)
    names = {k: (kind, cin) for kind, k, cin, _ in m.dec_plan + m.enc_plan}
    up = next(k for k, (kind, _) in names.items() if kind == "up3d")
    dn = next(k for k, (kind, _) in names.items() if kind == "down3d")
    cu, cd = names[up][1], names[dn][1]
    x = torch.randn(cu, 3, 8, 8, generator=g(6))
    want = wan_vae.upsample(P, up, x, True)
    got = m._up(x.permute(1, 2, 3, 0).contiguous().to(cuda), up, cu, True).permute(3, 0, 1, 2).cpu()
    assert got.shape == want.shape and rel(got, want) < 2e-3, (got.shape, want.shape, rel(got, want))
    x = torch.randn(cd, 5, 16, 16, generator=g(7))
    want = wan_vae.downsample(P, dn, x, True)
    got = m._down(x.permute(1, 2, 3, 0).contiguous().to(cuda), dn, cd, True).permute(3, 0, 1, 2).cpu()
    assert got.shape == want.shape and rel(got, want) < 2e-3, (got.shape, want.shape, rel(got, want))
    # a single frame: both temporal branches pass the frame through
    x1 = torch.randn(cu, 1, 8, 8, generator=g(8))
    got = m._up(x1.permute(1, 2, 3, 0).contiguous().to(cuda), up, cu, True).permute(3, 0, 1, 2).cpu()
    assert rel(got, wan_vae.upsample(P, up, x1, True)) < 2e-3


def test_mid_attention_and_resblock(cuda):
    from worldforge_b200 import vae as wvae
    cfg = wan_vae.VaeConfig(dim=8)
    P = wan_vae.init_params(cfg, 9)
    m = wvae.WfWanVAE(P, cuda, dim=8)
    x = torch.randn(32, 2, 6, 8, generator=g(10))
    xc = x.permute(1, 2, 3, 0).contiguous().to(cuda)
    got = m._attn(xc, "decoder.middle.1", 32).permute(3, 0, 1, 2).cpu()
    assert rel(got, wan_vae.attn_block(P, "decoder.middle.1", x)) < 2e-3
    got = m._res(xc, "decoder.middle.0", 32, 32)[0].permute(3, 0, 1, 2).cpu()
    assert rel(got, wan_vae.res_block(P, "decoder.middle.0", x)) < 2e-3


@pytest.mark.parametrize("dim,F_,H,W", [(8, 9, 64, 96), (8, 1, 32, 32), (96, 5, 32, 48)])
def test_encode_decode_match_oracle(cuda, dim, F_, H, W):
    from worldforge_b200 import vae as wvae
    cfg = wan_vae.VaeConfig(dim=dim)
    P = wan_vae.init_params(cfg, 11)
    m = wvae.WfWanVAE(P, cuda, dim=dim)
    video = torch.rand(3, F_, H, W, generator=g(12)) * 2 - 1
    want_mu = wan_vae.encode_mode(P, cfg, video)
    got_mu = m.encode(video.unsqueeze(0).to(cuda)).latent_dist.mode()[0].cpu()
    assert got_mu.shape == want_mu.shape
    assert rel(got_mu, want_mu) < 1e-2, rel(got_mu, want_mu)
    z = torch.randn(16, (F_ - 1) // 4 + 1, H // 8, W // 8, generator=g(13))
    want = wan_vae.decode(P, cfg, z)
    got = m.decode(z.unsqueeze(0).to(cuda))[0][0].cpu()
    assert got.shape == want.shape
    assert got.abs().max() <= 1.0
    assert rel(got, want) < 1e-2, rel(got, want)


def test_vae_rejects_cpu(cuda):
    from worldforge_b200 import lib, vae as wvae
    m = wvae.WfWanVAE.random_init(cuda, dim=8)
    with pytest.raises(lib.WfError):
        m.decode(torch.zeros(1, 16, 1, 4, 4))


@pytest.mark.parametrize("cuts", [True, False])
@pytest.mark.parametrize("world,F_,H,W", [(2, 5, 64, 48), (4, 9, 96, 32), (8, 5, 128, 32), (3, 1, 48, 32), (8, 33, 64, 32)])
def test_row_sharded_vae_is_bit_identical(cuda, world, F_, H, W, cuts):
    """enable_row_sharding: the stages of the sharded evaluation (row slabs with recomputed halo; the mid-block attention
    by frames), every rank's share stitched in rank order between stages, equal the single-GPU evaluation bit for bit -
    simulated here by running the ranks one after the other on one GPU (tools/ulysses_check.py does it with real ranks)."""
    from worldforge_b200 import vae as wvae
    m = wvae.WfWanVAE.random_init(cuda, dim=8, seed=3)
    m.level_cuts = cuts                                             # default True: one row stage per resolution level
    video = (torch.rand(1, 3, F_, H, W, generator=g(21)) * 2 - 1).to(cuda)
    z = torch.randn(1, 16, (F_ - 1) // 4 + 1, H // 8, W // 8, generator=g(22)).to(cuda)
    want_mu = m.encode(video).latent_dist.mode()
    want_dec = m.decode(z)[0]

    def run(which, full):
        stages = m.sharded_stages(which, full, world)
        assert len(stages) == (3 if not cuts else 6 if which == "enc" else 5)      # rows | frames (attention) | rows, + level cuts
        for stage in stages:
            parts, dim = [], None
            for r in range(world):
                part, dim, bounds = stage(r, full)
                assert part.shape[dim] == bounds[r + 1] - bounds[r]
                parts.append(part)
            full = wvae.assemble_rows(parts, dim)
        return full

    mu = wvae.lib.cl_to_planar(run("enc", video[0]).contiguous(), m.z_dim).unsqueeze(0)
    assert torch.equal(mu, want_mu)
    x = m._conv(wvae.lib.planar_to_cl(z[0].contiguous(), m.z_dim, round_tf32=True), "conv2", wvae.TAPS_1, m.z_dim, round_out=True)
    assert torch.equal(run("dec", x).unsqueeze(0), want_dec)


def test_diffusers_state_dict_names_load(cuda):
    """The engine's VAE built from a state dict under diffusers' ``AutoencoderKLWan`` names (what ``pipe.vae.state_dict()``
    hands over in infer_worldforge.py) equals the one built from ``WanVAE_`` names (the map itself is pinned against both
    reference classes in tests/test_oracle_pinning.py)."""
    import re
    from worldforge_b200 import vae as wvae
    a = wvae.WfWanVAE.random_init(cuda, dim=16, seed=9)
    sd = {}
    # invert the map on the vendored-name dict the random init produces
    vend = wvae.random_state_dict(dim=16, seed=9)
    res = {"residual.0": "norm1", "residual.2": "conv1", "residual.3": "norm2", "residual.6": "conv2", "shortcut": "conv_shortcut"}

    def resnet(rest):
        for k, v in res.items():
            if rest.startswith(k + "."):
                return v + rest[len(k):]
        return rest
    for k, v in vend.items():
        if k.startswith("conv1."):
            sd["quant_conv." + k[6:]] = v; continue
        if k.startswith("conv2."):
            sd["post_quant_conv." + k[6:]] = v; continue
        side, rest = k.split(".", 1)
        if rest.startswith("conv1."):
            nk = "conv_in." + rest[6:]
        elif rest.startswith("head.0."):
            nk = "norm_out." + rest[7:]
        elif rest.startswith("head.2."):
            nk = "conv_out." + rest[7:]
        elif rest.startswith("middle."):
            i, tail = rest[7:].split(".", 1)
            nk = f"mid_block.attentions.0.{tail}" if i == "1" else f"mid_block.resnets.{int(i) // 2}.{resnet(tail)}"
        elif rest.startswith("downsamples."):
            i, tail = rest[12:].split(".", 1)
            nk = f"down_blocks.{i}.{resnet(tail)}"
        else:
            m = re.match(r"upsamples\.(\d+)\.(.*)", rest)
            idx, tail = int(m.group(1)), m.group(2)
            blk, j = divmod(idx, 4)
            nk = f"up_blocks.{blk}.upsamplers.0.{tail}" if j == 3 else f"up_blocks.{blk}.resnets.{j}.{resnet(tail)}"
        sd[side + "." + nk] = v
    assert any(k.startswith("decoder.up_blocks.2.upsamplers.0.") for k in sd) and "quant_conv.weight" in sd
    b = wvae.WfWanVAE(sd, cuda, dim=16)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(1, 16, 2, 8, 12, generator=g).to(cuda)
    assert torch.equal(a.decode(z)[0], b.decode(z)[0])
    video = (torch.rand(1, 3, 5, 64, 96, generator=g) * 2 - 1).to(cuda)
    assert torch.equal(a.encode(video).latent_dist.mode(), b.encode(video).latent_dist.mode())
