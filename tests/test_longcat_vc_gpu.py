"""LongCat video continuation (generate_vc, pipeline_longcat_video.py:1010-1270) on the engine vs the
oracle: the caching pass over the clean condition frames, the forward of the noise frames against the KV cache, and the
continuation loop with and without the cache.  The oracle's KV path is pinned to the reference DiT in
tests/test_oracle_pinning.py::test_longcat_kv_cache_path_matches_reference_fixture."""
import pytest
import torch

from oracle import adapters, longcat_dit as old, longcat_sched as ols

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
DIT = dict(hidden_size=256, depth=2, num_heads=2, caption_channels=64, adaln_tembed_dim=32, frequency_embedding_size=32)


def g(seed):
    return torch.Generator().manual_seed(seed)


def _model(cuda):
    from worldforge_b200 import longcat
    ocfg, pcfg = old.LongCatConfig(**DIT), longcat.LongCatConfig(**DIT)
    P = old.init_params(ocfg, 3)
    return P, ocfg, longcat.WfLongCatTransformer.from_state_dict(P, pcfg, cuda)


rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_kv_cache_forward_matches_oracle(cuda):
    P, ocfg, m = _model(cuda)
    lat = torch.randn(2, 16, 5, 8, 12, generator=g(4))
    ctx = torch.randn(2, 1, 10, 64, generator=g(5)).to(BF)
    mask = torch.ones(2, 10, dtype=torch.int64); mask[0, 7:] = 0
    cond, noise = lat[:1, :, :2], lat[:, :, 2:]
    empty = torch.zeros(1, 1, 10, 64, dtype=BF)
    out_c, cache = m(cond.to(cuda).to(BF), torch.zeros(1, 2, device=cuda, dtype=BF), empty.to(cuda), return_kv=True, skip_crs_attn=True)
    want_c, ocache = old.dit_forward(P, ocfg, cond[0].to(BF), torch.zeros(2), empty[0, 0], num_cond_latents=0, amp=True, return_kv=True,
                                     skip_crs_attn=True)
    assert rel(out_c[0].cpu(), want_c) < 8e-3
    assert set(cache) == {0, 1} and cache[0][0].shape == (1, 2 * 4 * 6, 256)
    ts = torch.full((2, 3), 600.0)
    out_n = m(noise.to(cuda).to(BF), ts.to(cuda).to(BF), ctx.to(cuda), encoder_attention_mask=mask.to(cuda), num_cond_latents=2,
              kv_cache_dict=cache)                                      # one cached sample shared by the batch of two (CFG)
    for s in range(2):
        want = old.dit_forward(P, ocfg, noise[s].to(BF), ts[s], ctx[s, 0][mask[s] != 0], num_cond_latents=2, amp=True, kv_cache_dict=ocache)
        assert rel(out_n[s].cpu(), want) < 8e-3, (s, rel(out_n[s].cpu(), want))
    # the cached form equals the joint forward with the condition frames at timestep 0 (clean frames get no cross-attention)
    joint = m(lat[:1].to(cuda).to(BF), torch.tensor([[0.0, 0.0, 600.0, 600.0, 600.0]], device=cuda).to(BF), ctx[:1].to(cuda),
              encoder_attention_mask=mask[:1].to(cuda), num_cond_latents=2)
    assert rel(out_n[0], joint[0][:, 2:]) < 4e-3
    off = m(cond.to(cuda).to(BF), torch.zeros(1, 2, device=cuda, dtype=BF), empty.to(cuda), return_kv=True, skip_crs_attn=True,
            offload_kv_cache=True)[1]
    assert not off[0][0].is_cuda and torch.equal(off[1][1], cache[1][1].cpu())


@pytest.mark.parametrize("use_kv_cache", [True, False])
def test_continuation_loop_matches_oracle(cuda, use_kv_cache):
    from worldforge_b200 import longcat_pipeline as lp
    P, ocfg, m = _model(cuda)
    lat0 = torch.randn(1, 16, 5, 8, 12, generator=g(21))
    pe = torch.randn(2, 1, 8, 64, generator=g(22)).to(BF)
    pm = torch.ones(2, 8, dtype=torch.int64); pm[0, 5:] = 0
    so, sw = ols.OracleEuler(1000, 1.0), lp.WfFlowMatchEulerScheduler(1000, 1.0)
    ts_o = ols.vc_timesteps(so, 12, enhance_hf=True)
    ts_w = lp.vc_timesteps(sw, 12, enhance_hf=True, device=cuda)
    assert torch.equal(ts_o, ts_w.cpu()) and len(ts_o) == 16             # 6 steps above t = 500 + the 10-step uniform tail
    ts_o, ts_w = ts_o[:5], ts_w[:5]
    want = ols.vc_loop(adapters.OracleLongCatDit(P, ocfg, amp=True), so, lat0.clone(), pe, pm, 2, ts_o, use_kv_cache=use_kv_cache)
    got = lp.vc_loop(m, sw, lat0.clone().to(cuda), pe.to(cuda), pm.to(cuda), 2, ts_w, use_kv_cache=use_kv_cache)
    assert got.shape == want.shape == lat0.shape
    assert torch.equal(got[:, :, :2].cpu(), lat0[:, :, :2])               # the condition latents come back untouched
    assert rel(got.cpu(), want) < 1e-2, rel(got.cpu(), want)
