"""The reference's GPU execution path - "PyTorch + flash-attn" - restated for bench.py's ``gpu_reference`` record.

BENCH / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the comparator BASELINE.json's metric names ("vs the
reference's PyTorch+flash-attn path on the same box"), kind ``port-gpu``.  /root/reference does not
exist on the GPU box and has no installable package, so the path is restated op for op from the vendored statement:

* DiT      ``WanModel.forward`` under ``torch.autocast(bf16)`` (wan/modules/model.py:493-582; blocks :278-317): every
           nn.Linear is a separate cuBLAS bf16 GEMM (q, k, v, o, cross q/k/v/k_img/v_img/o, ffn.0, ffn.2 - the context
           projections are recomputed in every forward, :218-224), LayerNorm / RMSNorm / modulation / residual as eager
           fp32 torch ops, RoPE in complex128 (:43-70), attention through ``flash_attn.flash_attn_varlen_func`` with
           the casts of ``flash_attention`` (wan/modules/attention.py:24-130);
* VAE      the chunked, cached evaluation of ``WanVAE_.encode/decode`` (wan/modules/vae.py:516-568) on cuDNN with its
           default tf32 convolutions: oracle/wan_vae_stream.py on the device;
* loop     oracle/pipeline.py (pinned bit for bit against ``WanImageToVideoPipeline.__call__``) over oracle/unipc.py,
           whose FLF channel scoring is OpenCV Farneback on the host, as in the reference.

The DiT here reads the ENGINE's weight tensors in place (bf16 matrices = what autocast feeds cuBLAS; fp32 norms), so
both arms run the same random-init model and 33 GB of weights exist once.  Numerically this is oracle/wan_dit.py with
``amp=True`` (same rounding points) - tests/test_gpu_reference_gpu.py checks that on a small model.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import wan_dit, wan_vae, wan_vae_stream

BF, F32 = torch.bfloat16, torch.float32


def flash_attention(q, k, v):
    """attention.py:24-130 for one sample without padding: q [Lq,n,d], k/v [Lk,n,d] in any float dtype; inputs that are not
    half precision are cast to bf16 (:55-56), the result is cast back to q's dtype (:130)."""
    from flash_attn import flash_attn_varlen_func
    out_dtype = q.dtype
    half = lambda t: t if t.dtype in (torch.float16, BF) else t.to(BF)
    q, k, v = half(q), half(k), half(v)
    q, k = q.to(v.dtype), k.to(v.dtype)
    cu = lambda n: torch.tensor([0, n], dtype=torch.int32, device=q.device)
    x = flash_attn_varlen_func(q=q, k=k, v=v, cu_seqlens_q=cu(q.shape[0]), cu_seqlens_k=cu(k.shape[0]),
                               max_seqlen_q=q.shape[0], max_seqlen_k=k.shape[0], dropout_p=0.0, softmax_scale=None,
                               causal=False, window_size=(-1, -1), deterministic=False)
    return x.type(out_dtype)


class RefGpuTransformer:
    """Quacks like ``pipe.transformer``; executes the reference's op sequence with torch / cuBLAS / flash-attn."""

    def __init__(self, engine_tr):
        self.t, self.cfg = engine_tr, engine_tr.cfg
        self.dtype = BF
        self.config = SimpleNamespace(patch_size=self.cfg.patch)
        self.calls = 0
        self._rope = {}

    @staticmethod
    def lin(x, w, b):
        """nn.Linear under autocast(bf16): bf16 operands, fp32 accumulation inside cuBLAS, bf16 result."""
        return F.linear(x.to(BF), w, b)

    def rope(self, x, grid):
        """rope_apply (model.py:43-70): complex128 rotation, result in fp32."""
        if grid not in self._rope:
            self._rope[grid] = wan_dit.rope_angles(128, grid).to(x.device).unsqueeze(1)
        L, n, d = x.shape
        xc = torch.view_as_complex(x.to(torch.float64).reshape(L, n, d // 2, 2))
        return torch.view_as_real(xc * self._rope[grid]).flatten(2).to(F32)

    def block(self, i, x, e0, grid, ctx):
        t, c = self.t, self.cfg
        b, D, n = t.blocks[i], c.dim, c.num_heads
        L = x.shape[0]
        e = t.mod_all[i] + e0[0]                                                     # [6, D] fp32 (:297-298)
        h = wan_dit.layer_norm(x, c.eps).to(F32) * (1 + e[1]) + e[0]
        q = wan_dit.rms_norm(self.lin(h, b.qkv_w[:D], b.qkv_b[:D]), b.norm_q, c.eps).view(L, n, 128)
        k = wan_dit.rms_norm(self.lin(h, b.qkv_w[D:2 * D], b.qkv_b[D:2 * D]), b.norm_k, c.eps).view(L, n, 128)
        v = self.lin(h, b.qkv_w[2 * D:], b.qkv_b[2 * D:]).view(L, n, 128)
        a = flash_attention(self.rope(q, grid), self.rope(k, grid), v)
        x = x + self.lin(a.flatten(1), b.o_w, b.o_b) * e[2]
        hq = wan_dit.layer_norm(x, c.eps, b.n3_w, b.n3_b)
        ctx_img, ctx_txt = ctx[:c.img_len], ctx[c.img_len:]
        q = wan_dit.rms_norm(self.lin(hq, b.cq_w, b.cq_b), b.cnorm_q, c.eps).view(L, n, 128)
        k = wan_dit.rms_norm(self.lin(ctx_txt, b.ckv_w[:D], b.ckv_b[:D]), b.cnorm_k, c.eps).view(-1, n, 128)
        v = self.lin(ctx_txt, b.ckv_w[D:], b.ckv_b[D:]).view(-1, n, 128)
        ki = wan_dit.rms_norm(self.lin(ctx_img, b.ckvi_w[:D], b.ckvi_b[:D]), b.cnorm_ki, c.eps).view(-1, n, 128)
        vi = self.lin(ctx_img, b.ckvi_w[D:], b.ckvi_b[D:]).view(-1, n, 128)
        a = flash_attention(q, k, v) + flash_attention(q, ki, vi)                    # x = x + img_x (:227)
        x = x + self.lin(a.flatten(1), b.co_w, b.co_b)
        h = wan_dit.layer_norm(x, c.eps).to(F32) * (1 + e[4]) + e[3]
        y = self.lin(F.gelu(self.lin(h, b.f0_w, b.f0_b), approximate="tanh"), b.f2_w, b.f2_b)
        return x + y * e[5]

    @torch.no_grad()
    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None,
                 attention_kwargs=None, return_dict=False):
        t, c = self.t, self.cfg
        self.calls += 1
        cols, grid = wan_dit.patchify(hidden_states[0].to(BF), c)
        tok = self.lin(cols, t.patch_w, t.patch_b)                                  # bf16 token stream (:534)
        s = wan_dit.sinusoid(c.freq_dim, timestep.reshape(-1)[:1].cpu()).to(F32).to(tok.device)
        e = F.linear(F.silu(F.linear(s, t.t0_w, t.t0_b)), t.t2_w, t.t2_b)            # fp32 island (:546-550)
        e0 = F.linear(F.silu(e), t.tp_w, t.tp_b).unflatten(1, (6, c.dim))
        txt = encoder_hidden_states[0]
        if txt.shape[0] < c.text_len:
            txt = torch.cat([txt, txt.new_zeros(c.text_len - txt.shape[0], txt.shape[1])])
        txt = self.lin(F.gelu(self.lin(txt, t.txt0_w, t.txt0_b), approximate="tanh"), t.txt2_w, t.txt2_b)
        im = F.layer_norm(encoder_hidden_states_image[0].to(F32), (c.img_dim,), t.img_n0_w, t.img_n0_b, 1e-5)
        im = self.lin(F.gelu(self.lin(im, t.img1_w, t.img1_b)), t.img3_w, t.img3_b)
        im = F.layer_norm(im.to(F32), (c.dim,), t.img_n4_w, t.img_n4_b, 1e-5)
        ctx = torch.cat([im, txt.to(F32)], dim=0)
        for i in range(c.num_layers):
            tok = self.block(i, tok, e0, grid, ctx)
        eh = t.head_mod + e                                                          # [2, D]
        h = wan_dit.layer_norm(tok, c.eps).to(F32) * (1 + eh[1]) + eh[0]
        y = F.linear(h, t.head_w, t.head_b)                                          # fp32 (:344-346)
        return (wan_dit.unpatchify(y, c, grid).unsqueeze(0).to(self.dtype),)


class RefGpuVAE:
    """Quacks like ``pipe.vae``: the chunked, cached reference schedule on cuDNN (tf32 convolutions by default)."""

    def __init__(self, params, cfg: wan_vae.VaeConfig, device):
        self.dev = torch.device(device)
        self.P, self.cfg = {k: v.to(self.dev) for k, v in params.items()}, cfg
        self.dtype = F32
        self.config = SimpleNamespace(z_dim=cfg.z_dim, latents_mean=list(wan_vae.LATENTS_MEAN[:cfg.z_dim]),
                                      latents_std=list(wan_vae.LATENTS_STD[:cfg.z_dim]))
        self.temperal_downsample = list(cfg.temporal_downsample)

    @torch.no_grad()
    def encode(self, x):
        mu = wan_vae_stream.encode_mode(self.P, self.cfg, x[0].to(self.dev))
        return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: mu.unsqueeze(0)))

    @torch.no_grad()
    def decode(self, z, return_dict=False):
        return (wan_vae_stream.decode(self.P, self.cfg, z[0].to(self.dev)).unsqueeze(0),)
