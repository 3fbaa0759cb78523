"""bench.py's ``gpu_reference`` comparator (oracle/gpu_path.py: the reference's PyTorch + flash-attn execution path) is
the model it claims to be: on a small network it agrees with the CPU oracle (wan_dit ``amp=True`` / wan_vae) to the
bf16 / tf32 floor, i.e. both bench arms compute the same function of the same weights."""
import pytest
import torch

from oracle import gpu_path, wan_dit, wan_vae

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_reference_gpu_dit_is_the_oracle_forward(cuda):
    from worldforge_b200.transformer import WanDitConfig, WfWanTransformer
    kw = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16, img_dim=64, img_len=5, freq_dim=32)
    cfg = wan_dit.DitConfig(**kw)
    P = wan_dit.init_params(cfg, 7)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 36, 3, 8, 12, generator=g).to(torch.bfloat16)
    ctx = torch.randn(1, 16, 64, generator=g).to(torch.bfloat16)
    clip = torch.randn(1, 5, 64, generator=g).to(torch.bfloat16)
    t = torch.tensor([737])
    model = wan_dit.dit_forward(P, cfg, x[0], t, ctx[0], clip[0], amp=True).to(torch.bfloat16).float()
    truth = wan_dit.dit_forward(P, cfg, x[0], t, ctx[0], clip[0], amp=False)
    eng = WfWanTransformer.from_state_dict(P, WanDitConfig(**kw), cuda)
    ref = gpu_path.RefGpuTransformer(eng)
    got = ref(x.to(cuda), t.to(cuda), ctx.to(cuda), clip.to(cuda))[0][0].float().cpu()
    ours = eng(x.to(cuda), t.to(cuda), ctx.to(cuda), clip.to(cuda))[0][0].float().cpu()
    e_model, e_ref, e_ours = rel(model, truth), rel(got, truth), rel(ours, truth)
    print(f"\n[floor] comparator DiT (cuBLAS + flash-attn): vs fp32 {e_ref:.3e}; oracle amp {e_model:.3e}; engine {e_ours:.3e}")
    assert e_ref <= 1.1 * e_model and e_ours <= 1.1 * e_model


def test_reference_gpu_vae_is_the_oracle_vae(cuda):
    cfg = wan_vae.VaeConfig(dim=16)
    P = wan_vae.init_params(cfg, 5)
    g = torch.Generator().manual_seed(1)
    video = torch.rand(1, 3, 9, 32, 48, generator=g) * 2 - 1
    z = torch.randn(1, 16, 3, 4, 6, generator=g)
    vae = gpu_path.RefGpuVAE(P, cfg, cuda)
    mu = vae.encode(video.to(cuda)).latent_dist.mode()[0].cpu()
    dec = vae.decode(z.to(cuda))[0][0].cpu()
    assert rel(mu, wan_vae.encode_mode(P, cfg, video[0])) < 3e-3          # cuDNN tf32 convolutions against fp32
    assert rel(dec, wan_vae.decode(P, cfg, z[0])) < 3e-3
