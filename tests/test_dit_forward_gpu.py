"""Whole-forward parity of the CUDA DiT with the oracle (tiny widths, same topology and head_dim)."""
import pytest
import torch

from oracle import wan_dit

pytestmark = pytest.mark.gpu


def tiny_cfg(layers=2):
    return wan_dit.DitConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=layers, text_dim=64, text_len=16,
                             img_dim=64, img_len=5, freq_dim=32)


def product_cfg(o):
    from worldforge_b200.transformer import WanDitConfig
    return WanDitConfig(dim=o.dim, ffn_dim=o.ffn_dim, num_heads=o.num_heads, num_layers=o.num_layers, in_dim=o.in_dim,
                        out_dim=o.out_dim, freq_dim=o.freq_dim, text_dim=o.text_dim, text_len=o.text_len,
                        img_dim=o.img_dim, img_len=o.img_len)


def inputs(cfg, grid, seed=0):
    g = torch.Generator().manual_seed(seed)
    f, h, w = grid
    x = torch.randn(1, 36, f, h, w, generator=g).to(torch.bfloat16)
    ctx = torch.randn(1, cfg.text_len, cfg.text_dim, generator=g).to(torch.bfloat16)
    clip = torch.randn(1, cfg.img_len, cfg.img_dim, generator=g).to(torch.bfloat16)
    return x, ctx, clip


@pytest.mark.parametrize("grid,t", [((3, 8, 12), 737), ((2, 10, 18), 12)])
def test_forward_matches_oracle(cuda, grid, t):
    from worldforge_b200.transformer import WfWanTransformer
    cfg = tiny_cfg()
    P = wan_dit.init_params(cfg, 7)
    x, ctx, clip = inputs(cfg, grid)
    want = wan_dit.dit_forward(P, cfg, x[0], torch.tensor([t]), ctx[0], clip[0], amp=True)
    m = WfWanTransformer.from_state_dict(P, product_cfg(cfg), cuda)
    got = m(x.to(cuda), torch.tensor([t], device=cuda), ctx.to(cuda), clip.to(cuda), return_dict=False)[0]
    assert got.dtype == torch.bfloat16 and got.shape == (1, 16) + tuple(x.shape[2:])
    got = got[0].float().cpu()
    rel = ((got - want).norm() / want.norm()).item()
    # oracle-vs-reference spread under bf16 autocast is ~1e-3 (see tests/test_oracle_pinning.py); the
    # output itself is rounded to bf16 (2^-9 relative per element, ~1.1e-3 rms)
    assert rel < 4e-3, rel
    # second call with the same embeddings hits the context K/V cache and must give the same answer
    got2 = m(x.to(cuda), torch.tensor([t], device=cuda), ctx.to(cuda), clip.to(cuda), return_dict=False)[0]
    assert torch.equal(got2[0].float().cpu(), got)


def test_diffusers_key_names(cuda):
    """The same weights under diffusers' WanTransformer3DModel names load to the same model."""
    from worldforge_b200 import transformer as T
    cfg = tiny_cfg(layers=1)
    P = wan_dit.init_params(cfg, 9)
    inv_top = {v: k for k, v in T._DIFFUSERS_TOP.items()}
    inv_blk = {v: k for k, v in T._DIFFUSERS_BLOCK.items()}
    sd = {}
    for k, v in P.items():
        if k == "head.modulation":
            sd["scale_shift_table"] = v; continue
        parts = k.split(".")
        if parts[0] == "blocks":
            rest = ".".join(parts[2:])
            if rest == "modulation":
                sd[f"blocks.{parts[1]}.scale_shift_table"] = v; continue
            for src, dst in inv_blk.items():
                if rest.startswith(src + "."):
                    sd[f"blocks.{parts[1]}.{dst}{rest[len(src):]}"] = v; break
            else:
                sd[k] = v
            continue
        for src, dst in inv_top.items():
            if k.startswith(src + "."):
                sd[dst + k[len(src):]] = v; break
        else:
            sd[k] = v
    assert any(".attn1.to_q." in k for k in sd)
    x, ctx, clip = inputs(cfg, (2, 8, 8))
    a = T.WfWanTransformer.from_state_dict(P, product_cfg(cfg), cuda)
    b = T.WfWanTransformer.from_state_dict(sd, product_cfg(cfg), cuda)
    args = (x.to(cuda), torch.tensor([500], device=cuda), ctx.to(cuda), clip.to(cuda))
    assert torch.equal(a(*args)[0], b(*args)[0])
