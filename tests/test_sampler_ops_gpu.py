"""Parity of the fused WorldForge sampler kernels with the op-by-op torch evaluation the
reference performs (oracle/unipc.py, oracle/pipeline.py).  Integer-exact: bit-for-bit."""
import itertools

import numpy as np
import pytest
import torch

from oracle import flf as oflf
from oracle import unipc

pytestmark = pytest.mark.gpu
SHAPE = (1, 16, 3, 12, 20)
DT = {0: torch.float32, 1: torch.bfloat16}


def rnd(shape, dt, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * 1.7).to(dt)


def same(a_gpu, b_cpu):
    assert a_gpu.dtype == b_cpu.dtype, (a_gpu.dtype, b_cpu.dtype)
    a = a_gpu.cpu()
    if not torch.equal(a, b_cpu):
        d = (a.float() - b_cpu.float()).abs()
        raise AssertionError(f"mismatch: {int((d > 0).sum())} of {d.numel()} elements, max {d.max().item()}")


@pytest.mark.parametrize("bf", [0, 1])
def test_cfg_combine(cuda, bf):
    from worldforge_b200 import lib
    c, u = rnd(SHAPE, DT[bf], 1), rnd(SHAPE, DT[bf], 2)
    want = c + 4.0 * (c - u)
    same(lib.cfg_combine(c.to(cuda), u.to(cuda), 4.0), want)


@pytest.mark.parametrize("xb,vb", list(itertools.product([0, 1], [0, 1])))
def test_x0_convert(cuda, xb, vb):
    from worldforge_b200 import lib
    x, v = rnd(SHAPE, DT[xb], 3), rnd(SHAPE, DT[vb], 4)
    sigma = torch.tensor(0.8996, dtype=torch.float32)
    want = x - sigma * v
    same(lib.x0_convert(x.to(cuda), v.to(cuda), float(sigma)), want)


@pytest.mark.parametrize("xb,m0b,m1b,order", [(0, 0, 0, 1), (1, 1, 1, 1), (1, 1, 0, 2), (1, 1, 1, 2), (0, 0, 0, 2),
                                              (0, 1, 0, 2), (1, 0, 1, 2)])
def test_unip_update_matches_oracle_scheduler(cuda, xb, m0b, m1b, order):
    """Drive the oracle scheduler's predictor and the kernel with the same state."""
    from worldforge_b200 import lib, scheduler as wsched
    s = unipc.OracleUniPC(flow_shift=3.0)
    s.set_timesteps(10)
    s._step_index = 4
    x, m0, m1 = rnd(SHAPE, DT[xb], 5), rnd(SHAPE, DT[m0b], 6), rnd(SHAPE, DT[m1b], 7)
    s.model_outputs = [m1, m0]
    want = s.multistep_uni_p_bh_update(model_output=None, sample=x, order=order)
    co = wsched.unip_coefficients(s.sigmas, None, 4, order, resampling=False)
    got = lib.unip_update(x.to(cuda), m0.to(cuda), m1.to(cuda) if order == 2 else None, order, *co)
    same(got, want)


@pytest.mark.parametrize("bf", [0, 1])
def test_renoise(cuda, bf):
    from worldforge_b200 import lib
    x0 = rnd(SHAPE, DT[bf], 8)
    noise = rnd(SHAPE, torch.float32, 9)
    sig = torch.tensor([0.8996], dtype=torch.float32).to(DT[bf]).view(1, 1, 1, 1, 1)
    want = (1 - sig) * x0 + sig * noise
    got = lib.renoise(x0.to(cuda), noise.to(cuda), float((1 - sig).float()), float(sig.float()))
    same(got, want)


@pytest.mark.parametrize("bf", [0, 1])
def test_dsg(cuda, bf):
    from worldforge_b200 import lib
    g = rnd(SHAPE, DT[bf], 10)
    w = (g.float() * 0.8 + rnd(SHAPE, torch.float32, 11) * 0.5).to(DT[bf])
    dims = list(range(1, g.dim()))
    dot = torch.sum(g * w, dim=dims, keepdim=True)
    ng = torch.sqrt(torch.sum(g ** 2, dim=dims, keepdim=True))
    nw = torch.sqrt(torch.sum(w ** 2, dim=dims, keepdim=True))
    cos = dot / (ng * nw + 1e-8)
    sin = torch.sin(torch.acos(torch.clamp(cos, -1.0, 1.0)))
    ratio = ng / (nw + 1e-8)
    want = g + 4.0 * sin * (g - (ratio * cos) * w)
    stats = torch.zeros(3, device=cuda)
    got = lib.dsg(g.to(cuda), w.to(cuda), 4.0, stats)
    st = stats.cpu()
    # the three scalars come from fp32 reductions whose summation order differs from torch's:
    # they must agree to within one ulp of the tensor dtype, and when they agree exactly so must the output
    ulp = 2.0 ** -7 if bf else 2.0 ** -20
    assert abs(st[0] - cos.float().item()) <= ulp * abs(cos.float().item())
    assert abs(st[1] - sin.float().item()) <= ulp * max(abs(sin.float().item()), 1e-3)
    assert abs(st[2] - ratio.float().item()) <= ulp * abs(ratio.float().item())
    if st[0] == cos.float().item() and st[1] == sin.float().item() and st[2] == ratio.float().item():
        if bf:
            same(got, want)
        else:
            torch.testing.assert_close(got.cpu(), want, rtol=1e-6, atol=1e-6)
    else:
        torch.testing.assert_close(got.cpu().float(), want.float(), rtol=2e-2, atol=2e-2)


def test_flf_blend(cuda):
    from worldforge_b200 import lib
    g = torch.Generator().manual_seed(12)
    dec = torch.rand(1, 3, 5, 16, 24, generator=g) * 2 - 1
    ref = torch.rand(1, 3, 5, 16, 24, generator=g)
    m = torch.rand(1, 1, 5, 16, 24, generator=g)
    m[:, :, 0] = 1.0
    r = 2.0 * ref - 1.0
    mm = m.repeat(1, 3, 1, 1, 1)
    want = r * mm + dec * (1 - mm)
    same(lib.flf_blend(dec.to(cuda), ref.to(cuda), m.to(cuda)), want)


@pytest.mark.parametrize("bf", [0, 1])
def test_latent_denorm_norm_replace(cuda, bf):
    from worldforge_b200 import lib, scheduler as wsched
    from oracle import wan_vae
    dt = DT[bf]
    x0 = rnd(SHAPE, dt, 13)
    mean = torch.tensor(wan_vae.LATENTS_MEAN).view(1, 16, 1, 1, 1).to(dt)
    inv_std = 1.0 / torch.tensor(wan_vae.LATENTS_STD).view(1, 16, 1, 1, 1).to(dt)
    want = (x0 / inv_std + mean).to(torch.float32)
    mh, sh = wsched.latent_stats(wan_vae.LATENTS_MEAN, wan_vae.LATENTS_STD, dt)
    same(lib.latent_denorm(x0.to(cuda), mh, sh), want)
    enc = rnd(SHAPE, torch.float32, 14)
    e = (enc - mean) * inv_std
    for c in (2, 7, 15):
        e[:, c] = x0[:, c]
    want2 = e.to(dt)
    same(lib.latent_norm_replace(enc.to(cuda), x0.to(cuda), mh, sh, [2, 7, 15]), want2)


@pytest.mark.parametrize("bf", [0, 1])
def test_quantise_u8(cuda, bf):
    from worldforge_b200 import lib
    x = rnd(SHAPE, DT[bf], 15)
    want = oflf.quantise_u8(x)
    got = lib.quantise_u8(x.to(cuda)).cpu().numpy()[0]
    assert np.array_equal(got, want)


def test_rejects_cpu_tensors_and_bad_sizes(cuda):
    from worldforge_b200 import lib
    with pytest.raises(lib.WfError):
        lib.cfg_combine(torch.zeros(8), torch.zeros(8), 1.0)
    with pytest.raises(AssertionError):
        lib.cfg_combine(torch.zeros(6, device=cuda), torch.zeros(6, device=cuda), 1.0)
