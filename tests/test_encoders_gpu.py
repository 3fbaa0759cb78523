"""The once-per-video encoders on the engine against oracle/encoders.py, which reproduces the
reference's bf16 T5Encoder bit for bit and its fp32 T5 / CLIP classes to 1e-6 (tests/test_oracle_pinning.py)."""
import pytest
import torch

from oracle import encoders as oe

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
T5 = dict(vocab=50, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2, num_buckets=32, shared_pos=False)
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_t5_kernels(cuda):
    from worldforge_b200 import encoders, lib
    g = torch.Generator().manual_seed(0)
    L, H = 40, 3
    qkv = (torch.randn(L, 3 * H * 64, generator=g)).to(BF)
    emb = (0.5 * torch.randn(32, H, generator=g)).to(BF)
    bucket = encoders.relative_position_bucket(L)
    idx = torch.arange(L).unsqueeze(0) - torch.arange(L).unsqueeze(1)
    assert torch.equal(bucket[(idx + L - 1)].long(), oe.relative_position_bucket(idx))
    q, k, v = (qkv[:, i * H * 64:(i + 1) * H * 64].reshape(L, H, 64) for i in range(3))
    n_valid = 29
    att = torch.einsum("inc,jnc->nij", q.float(), k.float()).to(BF) + emb[oe.relative_position_bucket(idx)].permute(2, 0, 1)
    mask = torch.arange(L) < n_valid
    att = att.masked_fill(~mask.view(1, 1, -1), torch.finfo(BF).min)
    want = torch.einsum("nij,jnc->inc", torch.softmax(att.float(), -1).to(BF).float(), v.float()).to(BF).reshape(L, H * 64)
    d = qkv.to(cuda)
    out = torch.empty(L, H * 64, dtype=BF, device=cuda)
    lib.attention_small(d[:, :H * 64], d[:, H * 64:2 * H * 64], d[:, 2 * H * 64:], out, H, 64, mode=0, n_valid=n_valid,
                        bias_emb=emb.to(cuda), bias_bucket=bucket.to(cuda))
    dd = (out.cpu().float() - want.float()).abs()
    assert dd.max() <= 2 ** -6 and (dd > 0).float().mean() < 0.05, (dd.max().item(), (dd > 0).float().mean().item())
    gf = torch.randn(37, 512, generator=g).to(BF)
    ff = torch.empty(37, 256, dtype=BF, device=cuda)
    lib.geglu_bf16(gf.to(cuda), ff)
    wantg = gf[:, 256:] * oe.gelu_tanh_expr(gf[:, :256])
    dg = (ff.cpu().float() - wantg.float()).abs()
    # tanhf (device) vs torch's CPU tanh may round a bf16 value the other way; where 1 + tanh is a few bf16 steps from zero
    # (very negative gate) that is a whole step of the factor, i.e. |fc1 * 0.5 * gate| * 2^-8 in absolute terms
    bound = wantg.float().abs() * 2 ** -7 + (gf[:, 256:].float() * gf[:, :256].float()).abs() * 2 ** -8 + 1e-6
    assert (dg <= bound).all() and (dg > 0).float().mean() < 0.02, ((dg - bound).max().item(), (dg > 0).float().mean().item())


def test_t5_encoder_matches_oracle(cuda):
    from worldforge_b200 import encoders
    P = oe.init_params(oe.t5_shapes(**T5), 3)
    for i in range(2):                                  # T5 has no softmax scale: keep the random-weight logits O(1) so that one
        P[f"blocks.{i}.attn.q.weight"] *= 0.25          # flipped bf16 rounding of a score is not amplified by a saturated softmax
    m = encoders.WfT5Encoder(P, cuda, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 50, (2, 24), generator=g)
    mask = torch.ones(2, 24, dtype=torch.long); mask[0, 17:] = 0
    got = m(ids.to(cuda), mask.to(cuda)).last_hidden_state
    for s in range(2):
        want = oe.t5_encoder(P, ids[s], mask[s], 2, 2, amp=True)
        exact = oe.t5_encoder(P, ids[s], mask[s], 2, 2, amp=False)
        n = int(mask[s].sum())
        floor = rel(want[:n], exact[:n])                # what the reference's own bf16 module loses against fp32
        e_fp32, e_model = rel(got[s, :n].cpu(), exact[:n]), rel(got[s, :n].cpu(), want[:n])
        print(f"\n[floor] T5 encoder sample {s}: bf16-module-vs-fp32 {floor:.3e}  engine-vs-fp32 {e_fp32:.3e}  engine-vs-bf16-module {e_model:.3e}")
        assert e_fp32 <= 1.25 * floor + 1e-3 and e_model <= 1.5 * floor + 1e-3, (s, floor, e_fp32, e_model)
    pe = encoders.t5_prompt_embeds(m, ids.to(cuda), mask.to(cuda), max_sequence_length=32)
    assert pe.shape == (2, 32, 128) and not pe[0, 17:].any() and torch.equal(pe[1, :24], got[1])
    # transformers' UMT5EncoderModel parameter names load to the same network
    inv = {"attn.q": "layer.0.SelfAttention.q", "attn.k": "layer.0.SelfAttention.k", "attn.v": "layer.0.SelfAttention.v",
           "attn.o": "layer.0.SelfAttention.o", "pos_embedding.embedding": "layer.0.SelfAttention.relative_attention_bias",
           "norm1": "layer.0.layer_norm", "norm2": "layer.1.layer_norm", "ffn.gate.0": "layer.1.DenseReluDense.wi_0",
           "ffn.fc1": "layer.1.DenseReluDense.wi_1", "ffn.fc2": "layer.1.DenseReluDense.wo"}
    hf = {"shared.weight": P["token_embedding.weight"], "encoder.final_layer_norm.weight": P["norm.weight"]}
    for k, v in P.items():
        if k.startswith("blocks."):
            _, i, rest = k.split(".", 2)
            mod, leaf = rest.rsplit(".", 1)
            hf[f"encoder.block.{i}.{inv[mod]}.{leaf}"] = v
    m2 = encoders.WfT5Encoder(hf, cuda, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2)
    assert torch.equal(m2(ids.to(cuda), mask.to(cuda)).last_hidden_state, got)


def test_clip_vision_encoder_matches_oracle(cuda):
    from worldforge_b200 import encoders
    PC = oe.init_params(oe.clip_shapes(28, 14, 160, 4, 3), 5)
    m = encoders.WfCLIPVisionEncoder(PC, cuda, image_size=28, patch_size=14, dim=160, mlp_ratio=4, num_heads=2, num_layers=3)
    g = torch.Generator().manual_seed(2)
    img = torch.randn(2, 3, 28, 28, generator=g)
    got = m(img.to(cuda)).hidden_states[-2]
    assert got.shape == (2, 5, 160)
    for s in range(2):
        want = oe.clip_visual(PC, img[s], 14, 2, 3, amp=True)
        assert rel(got[s].cpu(), want) < 6e-3, rel(got[s].cpu(), want)
