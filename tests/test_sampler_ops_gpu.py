"""Parity of the fused WorldForge sampler kernels with the op-by-op torch evaluation the
reference performs (oracle/unipc.py, oracle/pipeline.py), bit-for-bit.

The expected values are the reference's own torch expressions evaluated ON THE GPU with the operands
placed where the reference places them (``scheduler.sigmas`` on the host, ``resample_sigmas`` on the
device): torch's CUDA kernels keep a host-side scalar operand in fp32 and cast a device-side 0-dim
tensor to the op's dtype, which its CPU kernels do not, so this - not a CPU evaluation - is what the
reference computes on a GPU."""
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
    c, u = rnd(SHAPE, DT[bf], 1).to(cuda), rnd(SHAPE, DT[bf], 2).to(cuda)
    want = c + 4.0 * (c - u)
    same(lib.cfg_combine(c, u, 4.0), want.cpu())


@pytest.mark.parametrize("xb,vb", list(itertools.product([0, 1], [0, 1])))
def test_x0_convert(cuda, xb, vb):
    from worldforge_b200 import lib
    x, v = rnd(SHAPE, DT[xb], 3).to(cuda), rnd(SHAPE, DT[vb], 4).to(cuda)
    sigma = torch.tensor(0.8996, dtype=torch.float32)          # host scalar, like scheduler.sigmas[i]
    want = x - sigma * v
    same(lib.x0_convert(x, v, float(sigma)), want.cpu())


@pytest.mark.parametrize("resampling", [False, True])
@pytest.mark.parametrize("xb,m0b,m1b,order", [(0, 0, 0, 1), (1, 1, 1, 1), (1, 1, 0, 2), (1, 1, 1, 2), (0, 0, 0, 2),
                                              (0, 1, 0, 2), (1, 0, 1, 2)])
def test_scheduler_step_matches_oracle_on_device(cuda, xb, m0b, m1b, order, resampling):
    """x0 conversion + UniP predictor: the engine's scheduler against the oracle scheduler, both on the GPU."""
    from worldforge_b200 import scheduler as wsched
    x, m0, m1 = rnd(SHAPE, DT[xb], 5).to(cuda), rnd(SHAPE, DT[m0b], 6).to(cuda), rnd(SHAPE, DT[m1b], 7).to(cuda)
    v = rnd(SHAPE, torch.bfloat16, 8).to(cuda)
    outs = []
    for cls in (unipc.OracleUniPC, wsched.WfUniPCScheduler):
        s = cls(flow_shift=3.0)
        s.set_timesteps(10, device=cuda)
        s._step_index = 4
        s.is_resampling = resampling
        s.model_outputs = [m1, m0]
        upd = s.multistep_uni_p_bh_update(model_output=None, sample=x, order=order)
        x0 = s.convert_model_output(v, sample=x)
        outs.append((upd, x0))
    (want_u, want_x0), (got_u, got_x0) = outs
    fp32_inputs = not (xb and m0b and (m1b or order == 1))
    if resampling and fp32_inputs:
        # device-side scalar math (log / expm1 on the GPU) may differ from the host's by an fp32 ulp
        assert got_u.dtype == want_u.dtype
        if got_u.dtype == torch.bfloat16:     # an fp32-ulp change of a coefficient can flip a final bf16 rounding
            d = (got_u.float() - want_u.float()).abs()
            assert (d <= want_u.float().abs() * 2.0 ** -7 + 1e-6).all() and (d > 0).float().mean() < 1e-2
        else:
            torch.testing.assert_close(got_u, want_u, rtol=2e-6, atol=2e-6)
    else:
        same(got_u, want_u.cpu())
    same(got_x0, want_x0.cpu())


@pytest.mark.parametrize("bf", [0, 1])
def test_renoise(cuda, bf):
    from worldforge_b200 import lib
    x0 = rnd(SHAPE, DT[bf], 8).to(cuda)
    noise = rnd(SHAPE, torch.float32, 9).to(cuda)
    sig = torch.tensor([0.8996], dtype=torch.float32).to(cuda).to(DT[bf]).view(1, 1, 1, 1, 1)
    want = (1 - sig) * x0 + sig * noise
    got = lib.renoise(x0, noise, float((1 - sig).float()), float(sig.float()))
    same(got, want.cpu())


@pytest.mark.parametrize("bf", [0, 1])
def test_dsg(cuda, bf):
    from worldforge_b200 import lib
    g = rnd(SHAPE, DT[bf], 10).to(cuda)
    w = (g.float() * 0.8 + rnd(SHAPE, torch.float32, 11).to(cuda) * 0.5).to(DT[bf])
    dims = list(range(1, g.dim()))
    dot = torch.sum(g * w, dim=dims, keepdim=True)
    ng = torch.sqrt(torch.sum(g ** 2, dim=dims, keepdim=True))
    nw = torch.sqrt(torch.sum(w ** 2, dim=dims, keepdim=True))
    cos = dot / (ng * nw + 1e-8)
    sin = torch.sin(torch.acos(torch.clamp(cos, -1.0, 1.0)))
    ratio = ng / (nw + 1e-8)
    want = g + 4.0 * sin * (g - (ratio * cos) * w)
    stats = torch.zeros(3, device=cuda)
    got = lib.dsg(g, w, 4.0, stats)
    st = stats.cpu()
    want, cos, sin, ratio = want.cpu(), cos.cpu(), sin.cpu(), ratio.cpu()
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
    dec = (torch.rand(1, 3, 5, 16, 24, generator=g) * 2 - 1).to(cuda)
    ref = torch.rand(1, 3, 5, 16, 24, generator=g).to(cuda)
    m = torch.rand(1, 1, 5, 16, 24, generator=g)
    m[:, :, 0] = 1.0
    m = m.to(cuda)
    r = 2.0 * ref - 1.0
    mm = m.repeat(1, 3, 1, 1, 1)
    want = r * mm + dec * (1 - mm)
    same(lib.flf_blend(dec, ref, m), want.cpu())


@pytest.mark.parametrize("bf", [0, 1])
def test_latent_denorm_norm_replace(cuda, bf):
    from worldforge_b200 import lib, scheduler as wsched
    from oracle import wan_vae
    dt = DT[bf]
    x0 = rnd(SHAPE, dt, 13).to(cuda)
    mean = torch.tensor(wan_vae.LATENTS_MEAN).view(1, 16, 1, 1, 1).to(cuda, dt)
    inv_std = 1.0 / torch.tensor(wan_vae.LATENTS_STD).view(1, 16, 1, 1, 1).to(cuda, dt)
    want = (x0 / inv_std + mean).to(torch.float32)
    mh, sh = wsched.latent_stats(wan_vae.LATENTS_MEAN, wan_vae.LATENTS_STD, dt)
    same(lib.latent_denorm(x0, mh, sh), want.cpu())
    enc = rnd(SHAPE, torch.float32, 14).to(cuda)
    e = (enc - mean) * inv_std
    for c in (2, 7, 15):
        e[:, c] = x0[:, c]
    want2 = e.to(dt)
    same(lib.latent_norm_replace(enc, x0, mh, sh, [2, 7, 15]), want2.cpu())


@pytest.mark.parametrize("bf", [0, 1])
def test_quantise_u8(cuda, bf):
    from worldforge_b200 import lib
    x = rnd(SHAPE, DT[bf], 15)
    want = oflf.quantise_u8(x.to(cuda))       # the reference normalises on the device, then copies to the host
    got = lib.quantise_u8(x.to(cuda)).cpu().numpy()[0]
    assert np.array_equal(got, want)
    assert np.array_equal(oflf.quantise_u8(x), want)   # and the CPU evaluation agrees (plain IEEE fp32 ops)


def test_rejects_cpu_tensors_and_bad_sizes(cuda):
    from worldforge_b200 import lib
    with pytest.raises(lib.WfError):
        lib.cfg_combine(torch.zeros(8), torch.zeros(8), 1.0)
    with pytest.raises(AssertionError):
        lib.cfg_combine(torch.zeros(6, device=cuda), torch.zeros(6, device=cuda), 1.0)
