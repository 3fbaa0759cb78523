"""Host-side logic of the engine that needs no GPU: schedules and solver coefficients, FLF policy and flow metrics,
weight re-packing for the fused convolution forms, RoPE tables, key-name mapping, synthetic inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import flf as oflf
from oracle import pipeline as opipe
from oracle import unipc, wan_dit, wan_vae
from worldforge_b200 import flf_select, scheduler as wsched, synth, transformer as wtr, vae as wvae


@pytest.mark.parametrize("n,shift", [(50, 3.0), (10, 5.0), (2, 3.0)])
def test_schedule_tables(n, shift):
    a, b = unipc.OracleUniPC(flow_shift=shift), wsched.WfUniPCScheduler(flow_shift=shift)
    a.set_timesteps(n); b.set_timesteps(n)
    assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas)
    assert torch.equal(a.resample_sigmas, b.resample_sigmas) and torch.equal(a.resample_timesteps, b.resample_timesteps)
    for i in (0, n - 1, n + 3):
        assert int(a.get_resample_timestep(i)) == int(b.get_resample_timestep(i))


def test_unip_coefficients_reproduce_the_oracle_update_on_cpu():
    """With fp32 tensors every coefficient convention collapses to plain fp32 arithmetic, so the closed form
    c_x*x - c_m0*m0 - c_res*0.5*(m1-m0)/r1 must reproduce the oracle's update."""
    s = unipc.OracleUniPC(flow_shift=3.0)
    s.set_timesteps(10)
    g = torch.Generator().manual_seed(0)
    x, m0, m1 = (torch.randn(1, 4, 2, 3, 3, generator=g) for _ in range(3))
    for idx, order in [(0, 1), (4, 2), (9, 1)]:
        s._step_index = idx
        s.model_outputs = [m1, m0]
        want = s.multistep_uni_p_bh_update(model_output=None, sample=x, order=order)
        c_x, c_m0, rk, c_res = wsched.unip_coefficients(s.sigmas, None, idx, order, False)
        got = c_x * x - c_m0 * m0
        if order == 2:
            got = got - c_res * (0.5 * ((m1 - m0) / rk))
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
    assert wsched.unip_coefficients(s.sigmas, None, 9, 1, False)[:2] == (0.0, -1.0)     # last step: x' = m0


def test_scheduler_state_machine_matches_oracle_without_tensors():
    """Order / index bookkeeping over a guided run (the pipeline pokes this state directly)."""
    class Fake:                                   # tensor stand-in: the bookkeeping must not depend on values
        dtype = torch.float32
    for cls in (unipc.OracleUniPC, wsched.WfUniPCScheduler):
        s = cls(flow_shift=3.0)
        s.set_timesteps(6)
        s.convert_model_output = lambda v, sample=None, **k: 0.0
        s.multistep_uni_p_bh_update = lambda model_output=None, sample=None, order=None, **k: order
        trace = []
        for i, t in enumerate(s.timesteps):
            for r in range(2 if i < 4 else 1):
                if r > 0:
                    s.set_resample_mode(True)
                    s._step_index -= 1
                    if s.lower_order_nums > 0 and s.last_lower_order_nums < 2:
                        s.lower_order_nums -= 1
                    s.this_order = s.last_this_order
                else:
                    s.set_resample_mode(False)
                out = s.step(0.0, t, 0.0, resampling=r > 0, is_resample_round=i < 4)
                trace.append((i, r, s._step_index, s.this_order, s.lower_order_nums, out.prev_sample))
            s.set_resample_mode(False)
        if cls is unipc.OracleUniPC:
            want = trace
    assert trace == want
    assert [t[3] for t in want if t[1] == 0] == [1, 2, 2, 2, 2, 1]       # first and last step are first order


def test_flf_policy_and_metrics_match_oracle():
    rng = np.random.RandomState(0)
    for step in (2, 5, 6, 10, 11, 30):
        for _ in range(20):
            sc = rng.rand(16).tolist()
            assert flf_select.selection_policy(sc, step) == oflf.policy(sc, step)
    tied = [0.5] * 16
    assert flf_select.selection_policy(tied, 20) == oflf.policy(tied, 20) and len(oflf.policy(tied, 20)) == 2
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(20, 2, 12, 16, generator=g) * 3, torch.randn(20, 2, 12, 16, generator=g) * 3
    assert flf_select.flow_similarity(a, b) == oflf.flow_similarity(a.unsqueeze(0), b.unsqueeze(0))


def test_farneback_front_end_matches_oracle():
    import cv2
    v = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(cv2.cvtColor(np.repeat(v[:, :, None], 3, axis=2), cv2.COLOR_RGB2GRAY), v)   # RGB2GRAY(v,v,v) == v
    u8 = (np.random.RandomState(2).rand(2, 4, 24, 32) * 255).astype(np.uint8)
    sel = flf_select.FlowChannelSelector(threads=2)
    got = sel._flows(u8)
    for c in range(2):
        assert torch.equal(got[c], oflf.farneback_flows(u8[c])[0])


def test_upsample_conv_as_four_parity_convs():
    """nearest-exact 2x upsample + 3x3 conv == for each output parity a 2x2-tap conv on the low-res input."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 5, 7, generator=g)           # [T, C, H, W]
    w = torch.randn(4, 6, 3, 3, generator=g)
    want = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest-exact"), w, padding=1)
    ws = wvae._w_upsample_parity(w)
    out = torch.zeros_like(want)
    for p in (0, 1):
        for q in (0, 1):
            taps = wvae._taps_upsample(p, q)
            wp = ws[p * 2 + q][:, :6].reshape(4, 4, 6)  # [tap, Co, Ci]
            for k, (_, dy, dx) in enumerate(taps):
                xs = F.pad(x, (1, 1, 1, 1))[:, :, 1 + dy:1 + dy + 5, 1 + dx:1 + dx + 7]
                out[:, :, p::2, q::2] += torch.einsum("oc,tchw->tohw", wp[k], xs)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)


def test_stride2_conv_as_space_to_depth_conv():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 4, 8, 12, generator=g)
    w = torch.randn(5, 4, 3, 3, generator=g)
    want = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, stride=2)
    ws = wvae._w_s2d(w).reshape(2, 2, 5, 16)            # [a, b, Co, (p,q,c)]
    s2d = x.reshape(2, 4, 4, 2, 6, 2).permute(0, 3, 5, 1, 2, 4).reshape(2, 16, 4, 6)   # channel = (p*2+q)*4 + c
    out = torch.zeros_like(want)
    sp = F.pad(s2d, (0, 1, 0, 1))
    for a in range(2):
        for b in range(2):
            out += torch.einsum("oc,tchw->tohw", ws[a, b], sp[:, :, a:a + 4, b:b + 6])
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)


def test_rope_table_and_key_names():
    grid = (3, 4, 5)
    ang = wan_dit.rope_angles(128, grid)
    tab = wtr.rope_table(grid)
    assert torch.equal(tab[..., 0], ang.real) and torch.equal(tab[..., 1], ang.imag)
    cfg = wan_dit.DitConfig(dim=128, ffn_dim=256, num_heads=1, num_layers=1, text_dim=32, text_len=8, img_dim=16, img_len=3, freq_dim=32)
    P = wan_dit.init_params(cfg, 0)
    assert wtr.to_vendored_names(P) is P
    diff = {"blocks.0.attn1.to_q.weight": 1, "blocks.0.norm2.weight": 2, "blocks.0.scale_shift_table": 3, "scale_shift_table": 4,
            "condition_embedder.time_proj.bias": 5, "proj_out.weight": 6, "blocks.0.attn2.add_k_proj.bias": 7,
            "blocks.0.ffn.net.0.proj.weight": 8, "patch_embedding.weight": 9}
    m = wtr.to_vendored_names(diff)
    assert m == {"blocks.0.self_attn.q.weight": 1, "blocks.0.norm3.weight": 2, "blocks.0.modulation": 3, "head.modulation": 4,
                 "time_projection.1.bias": 5, "head.head.weight": 6, "blocks.0.cross_attn.k_img.bias": 7,
                 "blocks.0.ffn.0.weight": 8, "patch_embedding.weight": 9}
    assert set(wan_dit.param_shapes(cfg)) == set(P)


def test_synthetic_inputs_and_soften_mask():
    inp = synth.make_inputs(9, 64, 96, text_len=8, text_dim=32, img_len=3, img_dim=16)
    assert inp.latents.shape == (1, 16, 3, 8, 12) and inp.condition.shape == (1, 20, 3, 8, 12)
    assert torch.equal(inp.condition[:, :4], opipe.first_frame_mask(9, 8, 12))
    assert inp.video_ref.shape == (1, 3, 9, 64, 96) and 0 <= inp.video_ref.min() and inp.video_ref.max() <= 1
    assert inp.mask.shape == (1, 1, 9, 64, 96) and (inp.mask[:, :, 0] == 1).all()
    m = np.zeros((2, 32, 40), np.float32); m[1, :, :20] = 1
    s = synth.soften_mask(m, 15, "sine")
    assert (s[0] == 0).all() and s[1, 0, 0] == 1 and 0 < s[1, 0, 19] < 0.2 and (s[1, :, 20:] == 0).all()
    assert np.all(np.diff(s[1, 0, :20]) <= 1e-6)          # ramps down towards the boundary
    with pytest.raises(ValueError):
        synth.soften_mask(m, 15, "nope")
    assert (torch.roll(inp.video_ref[0, :, 0], 2, dims=2) == inp.video_ref[0, :, 1]).all()     # 2 px per frame


def test_latent_stats_rounding():
    mh, sh = wsched.latent_stats(wan_vae.LATENTS_MEAN, wan_vae.LATENTS_STD, torch.bfloat16)
    assert mh[0] == float(torch.tensor(wan_vae.LATENTS_MEAN[0]).bfloat16())
    assert sh[3] == float((1.0 / torch.tensor(wan_vae.LATENTS_STD[3]).bfloat16()).float())
    m32, s32 = wsched.latent_stats(wan_vae.LATENTS_MEAN, wan_vae.LATENTS_STD, torch.float32)
    assert abs(s32[3] - 1 / 2.6558) < 1e-7


def test_longcat_host_logic_matches_oracle():
    from oracle import longcat_sched as ols
    from worldforge_b200 import longcat, longcat_pipeline as wlp
    from oracle import longcat_dit as old
    for n, d in [(50, False), (16, True), (4, True)]:
        assert torch.equal(wlp.timesteps_sigmas(n, d), ols.timesteps_sigmas(n, d))
        a, b = ols.OracleEuler(shift=1.0), wlp.WfFlowMatchEulerScheduler(shift=1.0)
        a.set_timesteps(n, sigmas=ols.timesteps_sigmas(n, d)); b.set_timesteps(n, sigmas=wlp.timesteps_sigmas(n, d))
        assert torch.equal(a.sigmas, b.sigmas) and torch.equal(a.timesteps, b.timesteps)
    rng = np.random.RandomState(1)
    for step in (2, 3, 4, 5, 6, 20):
        for dist in (False, True):
            for cap in (None, 1, 3):
                sc = rng.rand(16).tolist()
                assert wlp.selection_policy(sc, step, dist, cap) == ols.policy(sc, step, dist, cap)
    g = torch.Generator().manual_seed(2)
    x, y = torch.randn(8, 2, 12, 16, generator=g) * 3, torch.randn(8, 2, 12, 16, generator=g) * 3
    assert wlp.flow_similarity(x, y) == ols.flow_similarity(x.unsqueeze(0), y.unsqueeze(0))
    fr = old.rope_freqs(128, (2, 3, 4))
    tab = longcat.rope_table((2, 3, 4))
    assert torch.equal(tab[..., 0], fr.cos()[:, 0::2]) and torch.equal(tab[..., 1], fr.sin()[:, 0::2])
    assert longcat.LongCatConfig().ffn_dim == old.LongCatConfig().ffn_dim == 11008


def test_context_cache_is_keyed_by_content_not_by_address():
    """a cache keyed by (data_ptr, _version) alone returns the previous prompt's K|V when the
    allocator recycles the address.  Entries hold their sources; a fresh tensor hits only if its bytes are equal."""
    from worldforge_b200.ctx_cache import ContextCache
    c = ContextCache(2)
    a, img = torch.randn(4, 8), torch.randn(3, 8)
    assert c.get((a, img)) is None
    c.put((a, img), "A")
    assert c.get((a, img)) == "A" and c.hits_identity == 1
    assert c.get((a[:], img)) == "A" and c.hits_identity == 2          # a view of the same storage
    assert c.get((a.clone(), img.clone())) == "A" and c.hits_content == 1
    b = torch.randn(4, 8)
    assert c.get((b, img)) is None                                     # same shape, other prompt
    ptr = a.data_ptr()
    held_alive = c._entries[0][0][0]
    del a
    assert held_alive.data_ptr() == ptr                                # the entry keeps the storage from being recycled
    held_alive.add_(1.0)                                               # in-place write after caching: entry is stale
    assert c.get((held_alive, img)) is None
    c.put((held_alive, img), "A2")
    assert len(c) == 1 and c.get((held_alive, img)) == "A2"            # the stale entry was dropped
    c.put((b, img), "B"); c.put((torch.randn(4, 8), img), "C")
    assert len(c) == 2 and c.get((b, img)) == "B"                      # capacity 2: oldest evicted


def test_longcat_lora_surface_folds_and_restores():
    """load_lora / enable_loras / disable_all_loras of the reference DiT (longcat_video_dit.py:197-270; used at
    run_longcat_worldforge_single.py:213-214, 449-451, 490) on the engine's transformer: enabling folds exactly what
    ``merge_lora`` computes into the weights the decorated Linears own (fused qkv / kv / w1|w3 / adaLN rows included),
    two LoRAs add up, and disabling restores the original weights bit for bit.  (Plain torch on the host: the fold never
    touches a kernel.)"""
    import torch
    from oracle import longcat_dit as old
    from worldforge_b200 import longcat
    ocfg = old.LongCatConfig(hidden_size=128, depth=2, num_heads=1, caption_channels=32, adaln_tembed_dim=16, frequency_embedding_size=16)
    sd = old.init_params(ocfg, 3)
    cfg = longcat.LongCatConfig(hidden_size=128, depth=2, num_heads=1, caption_channels=32, adaln_tembed_dim=16, frequency_embedding_size=16)
    m = longcat.WfLongCatTransformer.from_state_dict(sd, cfg, "cpu")
    H, r = "___lorahyphen___", 4
    g = torch.Generator().manual_seed(11)
    C, Fd = cfg.hidden_size, cfg.ffn_dim

    def lora_for(mods, seed):
        g = torch.Generator().manual_seed(seed)
        out = {}
        for mod, (n_out, n_in, nsep) in mods.items():
            name = "lora" + H + mod.replace(".", H)
            out[name + ".lora_down.weight"] = torch.randn(nsep * r, n_in, generator=g) * 0.1
            if nsep == 1:
                out[name + ".lora_up.weight"] = torch.randn(n_out, r, generator=g) * 0.1
            else:
                for i in range(nsep):
                    out[f"{name}.lora_up.blocks.{i}.weight"] = torch.randn(n_out // nsep, r, generator=g) * 0.1
        return out
    la = lora_for({"blocks.0.attn.qkv": (3 * C, C, 3), "blocks.1.cross_attn.kv_linear": (2 * C, C, 2), "blocks.1.ffn.w3": (Fd, C, 1),
                   "blocks.0.adaLN_modulation.1": (6 * C, 16, 1), "final_layer.linear": (64, C, 1)}, 5)
    lb = lora_for({"blocks.0.attn.qkv": (3 * C, C, 3), "blocks.1.ffn.w1": (Fd, C, 1)}, 6)
    before = {k: v.clone() for k, v in (("qkv0", m.blocks[0].qkv_w), ("ckv1", m.blocks[1].ckv_w), ("w13_1", m.blocks[1].w13),
                                        ("ada", m.ada_w), ("final", m.final_w), ("proj0", m.blocks[0].proj_w))}
    m.load_lora(la, "a", multiplier=0.7, lora_network_dim=r, lora_network_alpha=2)
    m.load_lora(lb, "b", multiplier=1.0, lora_network_dim=r, lora_network_alpha=2)
    m.enable_loras(["a", "b"])
    assert m.active_loras == ["a", "b"]
    folded = longcat.merge_lora(longcat.merge_lora({k: v.to(torch.bfloat16).float() for k, v in sd.items()}, la, 0.7, r, 2), lb, 1.0, r, 2)
    bf = torch.bfloat16
    assert torch.equal(m.blocks[0].qkv_w, folded["blocks.0.attn.qkv.weight"].to(bf))
    assert torch.equal(m.blocks[1].ckv_w, folded["blocks.1.cross_attn.kv_linear.weight"].to(bf))
    assert torch.equal(m.blocks[1].w13, torch.cat([folded["blocks.1.ffn.w1.weight"], folded["blocks.1.ffn.w3.weight"]]).to(bf))
    assert torch.equal(m.ada_w[:6 * C], folded["blocks.0.adaLN_modulation.1.weight"].to(bf))
    assert torch.equal(m.ada_w[6 * C:], before["ada"][6 * C:])                        # untouched rows
    assert torch.equal(m.final_w, folded["final_layer.linear.weight"].to(bf).float())
    assert torch.equal(m.blocks[0].proj_w, before["proj0"]) and not torch.equal(m.blocks[0].qkv_w, before["qkv0"])
    m.enable_loras(["a"])                                                             # re-enabling starts from the originals
    only_a = longcat.merge_lora({k: v.to(bf).float() for k, v in sd.items()}, la, 0.7, r, 2)
    assert torch.equal(m.blocks[0].qkv_w, only_a["blocks.0.attn.qkv.weight"].to(bf)) and m.active_loras == ["a"]
    assert torch.equal(m.blocks[1].w13[:Fd], before["w13_1"][:Fd])
    m.disable_all_loras()
    assert m.active_loras == []
    for k, v in (("qkv0", m.blocks[0].qkv_w), ("ckv1", m.blocks[1].ckv_w), ("w13_1", m.blocks[1].w13), ("ada", m.ada_w),
                 ("final", m.final_w)):
        assert torch.equal(v, before[k]), k
    with pytest.raises(Exception):
        m.load_lora(lora_for({"blocks.0.attn.q_norm": (128, 128, 1)}, 7), "bad", lora_network_dim=r)
        m.enable_loras(["bad"])


def test_longcat_continuation_schedule_host():
    """generate_vc's timestep surgery (pipeline_longcat_video.py:1151-1164, enhance_hf): engine scheduler vs oracle scheduler on
    the host - the head of the standard schedule above t = 500, then ten uniform steps 500 -> 50, sigmas = t / 1000 + a final 0."""
    import torch
    from oracle import longcat_sched as ols
    from worldforge_b200 import longcat_pipeline as lp
    so, sw = ols.OracleEuler(1000, 1.0), lp.WfFlowMatchEulerScheduler(1000, 1.0)
    for n in (12, 50):
        a, b = ols.vc_timesteps(so, n, enhance_hf=True), lp.vc_timesteps(sw, n, enhance_hf=True)
        assert torch.equal(a, b) and torch.equal(so.sigmas, sw.sigmas)
        assert a[-10:].tolist() == [500.0, 450.0, 400.0, 350.0, 300.0, 250.0, 200.0, 150.0, 100.0, 50.0] and bool((a[:-10] > 500).all())
        assert float(sw.sigmas[-1]) == 0.0 and len(sw.sigmas) == len(b) + 1
    a, b = ols.vc_timesteps(so, 16, use_distill=True, enhance_hf=False), lp.vc_timesteps(sw, 16, use_distill=True, enhance_hf=False)
    assert torch.equal(a, b) and len(a) == 16
    with pytest.raises(Exception):
        lp.vc_timesteps(sw, 16, use_distill=True, enhance_hf=True)


def test_2d_context_parallel_split_is_the_reference_split():
    """ulysses.split_2d / gather_2d against the reference's own split_tensor_in_cp_2d (context_parallel_util.py:91-121) for every
    rank of the 1x2, 2x2 and 2x4 (cp = 8) layouts, and the layout rule get_optimal_split (:238-243) bench.py restates."""
    import importlib.util
    import os
    from worldforge_b200 import ulysses
    x = torch.arange(2 * 3 * 8 * 12 * 5).reshape(2, 3, 8, 12, 5)
    for split in ([1, 2], [2, 2], [2, 4], [4, 2]):
        parts = [ulysses.split_2d(x, (2, 3), split, r) for r in range(split[0] * split[1])]
        assert all(p.shape == (2, 3, 8 // split[0], 12 // split[1], 5) for p in parts)
        assert torch.equal(ulysses.gather_2d(parts, (2, 3), split), x)
        assert torch.equal(parts[1], x[:, :, :8 // split[0], 12 // split[1]:2 * (12 // split[1])])     # row-major: w inner
    with pytest.raises(RuntimeError):
        ulysses.split_2d(x, (2, 3), [3, 1], 0)
    optimal = lambda n: min(([i, n // i] for i in range(1, int(n ** 0.5) + 1) if n % i == 0), key=lambda f: abs(f[0] - f[1]))
    assert [optimal(n) for n in (1, 2, 4, 8)] == [[1, 1], [1, 2], [2, 2], [2, 4]]
    ref = "/root/reference/longcat_for_worldforge/longcat_video/context_parallel/context_parallel_util.py"
    if not os.path.exists(ref):
        return
    spec = importlib.util.spec_from_file_location("wf_ref_cp_util", ref)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert [m.get_optimal_split(n) for n in (1, 2, 4, 8)] == [optimal(n) for n in (1, 2, 4, 8)]
    for split in ([1, 2], [2, 2], [2, 4], [4, 2]):
        m.cp_size = split[0] * split[1]
        for r in range(m.cp_size):
            m.get_cp_rank = lambda r=r: r
            assert torch.equal(m.split_tensor_in_cp_2d(x, (2, 3), split), ulysses.split_2d(x, (2, 3), split, r)), (split, r)
