"""Oracle: LongCat-Video DiT forward (the i2v form WorldForge drives), restated functionally on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``longcat_for_worldforge/longcat_video/modules``: ``longcat_video_dit.py``
(LongCatVideoTransformer3DModel.forward :280-370, LongCatSingleStreamBlock.forward :68-121), ``attention.py``
(Attention.forward :107-147 with the condition / noise split :124-135, MultiHeadCrossAttention :211-276), ``blocks.py``
(TimestepEmbedder :173-198, CaptionEmbedder :206-218, FeedForwardSwiGLU :38-39, RMSNorm_FP32 :46-52, LayerNorm_FP32
:60-69, modulate_fp32 :120-128, FinalLayer_FP32 :147-156) and ``rope_3d.py`` (:63-119).

Two numeric modes, as in oracle/wan_dit.py:

* ``amp=False``: everything fp32 - pins the restatement against the imported reference module on the CPU.
* ``amp=True``: the dtype flow of the reference's GPU configuration: the module is cast to bf16
  (run_longcat_worldforge_single.py:207) and wraps its fp32 islands in ``amp.autocast('cuda', dtype=torch.float32)``
  (longcat_video_dit.py:82,101,114,312; blocks.py:152), i.e. inside them nn.Linear sees fp32 copies of the bf16
  weights.  Residual stream bf16 (rounded after every gated add, :103,:116), per-head RMSNorm with a bf16 gain, RoPE in
  fp32, SwiGLU in bf16, final projection in fp32.  NOTE: this relies on a torch whose CUDA autocast honours
  dtype=float32; torch 2.11 (this image) disables it with a warning, under which the reference's bf16 module raises a
  dtype error at the first fp32 island - the reference's own GPU path cannot run unmodified here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

BF16, F32 = torch.bfloat16, torch.float32


@dataclass
class LongCatConfig:
    in_channels: int = 16
    out_channels: int = 16
    hidden_size: int = 4096
    depth: int = 48
    num_heads: int = 32
    caption_channels: int = 4096
    mlp_ratio: int = 4
    adaln_tembed_dim: int = 512
    frequency_embedding_size: int = 256
    patch: Tuple[int, int, int] = (1, 2, 2)

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    @property
    def ffn_dim(self):                      # FeedForwardSwiGLU.__init__ (blocks.py:26-31)
        h = int(2 * int(self.hidden_size * self.mlp_ratio) / 3)
        return 256 * ((h + 255) // 256)


LONGCAT_13B = LongCatConfig()


def param_shapes(cfg: LongCatConfig) -> Dict[str, Tuple[int, ...]]:
    C, A, Fd = cfg.hidden_size, cfg.adaln_tembed_dim, cfg.ffn_dim
    pt, ph, pw = cfg.patch
    s = {
        "x_embedder.proj.weight": (C, cfg.in_channels, pt, ph, pw), "x_embedder.proj.bias": (C,),
        "t_embedder.mlp.0.weight": (A, cfg.frequency_embedding_size), "t_embedder.mlp.0.bias": (A,),
        "t_embedder.mlp.2.weight": (A, A), "t_embedder.mlp.2.bias": (A,),
        "y_embedder.y_proj.0.weight": (C, cfg.caption_channels), "y_embedder.y_proj.0.bias": (C,),
        "y_embedder.y_proj.2.weight": (C, C), "y_embedder.y_proj.2.bias": (C,),
        "final_layer.linear.weight": (pt * ph * pw * cfg.out_channels, C), "final_layer.linear.bias": (pt * ph * pw * cfg.out_channels,),
        "final_layer.adaLN_modulation.1.weight": (2 * C, A), "final_layer.adaLN_modulation.1.bias": (2 * C,),
    }
    for i in range(cfg.depth):
        b = f"blocks.{i}."
        s[b + "adaLN_modulation.1.weight"] = (6 * C, A); s[b + "adaLN_modulation.1.bias"] = (6 * C,)
        s[b + "attn.qkv.weight"] = (3 * C, C); s[b + "attn.qkv.bias"] = (3 * C,)
        s[b + "attn.q_norm.weight"] = (cfg.head_dim,); s[b + "attn.k_norm.weight"] = (cfg.head_dim,)
        s[b + "attn.proj.weight"] = (C, C); s[b + "attn.proj.bias"] = (C,)
        s[b + "cross_attn.q_linear.weight"] = (C, C); s[b + "cross_attn.q_linear.bias"] = (C,)
        s[b + "cross_attn.kv_linear.weight"] = (2 * C, C); s[b + "cross_attn.kv_linear.bias"] = (2 * C,)
        s[b + "cross_attn.proj.weight"] = (C, C); s[b + "cross_attn.proj.bias"] = (C,)
        s[b + "cross_attn.q_norm.weight"] = (cfg.head_dim,); s[b + "cross_attn.k_norm.weight"] = (cfg.head_dim,)
        s[b + "pre_crs_attn_norm.weight"] = (C,); s[b + "pre_crs_attn_norm.bias"] = (C,)
        s[b + "ffn.w1.weight"] = (Fd, C); s[b + "ffn.w2.weight"] = (C, Fd); s[b + "ffn.w3.weight"] = (Fd, C)
    return s


def init_params(cfg: LongCatConfig, seed: int = 2468) -> Dict[str, torch.Tensor]:
    """Random-init fp32 weights (N(0,0.02^2) matrices and biases, gains 1+N(0,0.05^2)); the bf16 model of amp mode is
    these values rounded to bf16."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if len(shape) == 1 and name.endswith("weight"):
            out[name] = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            out[name] = 0.02 * torch.randn(shape, generator=g)
    return out


def _rb(x):
    return x.to(BF16).to(F32)


def lin(x, w, b, amp):
    """nn.Linear of the bf16 module: bf16 operands, fp32 accumulation, bf16 result."""
    if not amp:
        return F.linear(x.to(F32), w, b)
    return F.linear(_rb(x), _rb(w), None if b is None else _rb(b)).to(BF16)


def lin_fp32_island(x, w, b, amp):
    """nn.Linear inside an autocast(float32) island: fp32 math on the (bf16-valued) weights."""
    if not amp:
        return F.linear(x.to(F32), w, b)
    return F.linear(x.to(F32), _rb(w), None if b is None else _rb(b))


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=F32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rope_freqs(head_dim: int, grid):
    """[T*H*W, head_dim] fp32 angles (rope_3d.py:63-95, no context-parallel split)."""
    T, H, W = grid
    d6 = head_dim // 6
    dims = (head_dim - 4 * d6, 2 * d6, 2 * d6)
    parts = []
    for n, dim in zip((T, H, W), dims):
        f = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        a = torch.arange(n, dtype=F32)[:, None] * f[None]
        parts.append(a.repeat_interleave(2, dim=-1))
    ft, fh, fw = parts
    full = torch.cat([ft[:, None, None, :].expand(T, H, W, -1), fh[None, :, None, :].expand(T, H, W, -1),
                      fw[None, None, :, :].expand(T, H, W, -1)], dim=-1)
    return full.reshape(T * H * W, head_dim)


def rotate_half(x):
    x1, x2 = x[..., 0::2], x[..., 1::2]
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def rope_apply(q, freqs):
    """q [N, H, D] -> fp32 rotation -> q.dtype (rope_3d.py:99-119)."""
    qf = q.float()
    cos, sin = freqs.cos()[:, None, :], freqs.sin()[:, None, :]
    return (qf * cos + rotate_half(qf) * sin).to(q.dtype)


def rms_head(x, w, amp):
    """RMSNorm_FP32 over head_dim (blocks.py:46-52); the gain is a bf16 parameter in the bf16 module."""
    xf = x.float()
    y = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(x.dtype)
    return y * (w.to(BF16) if amp else w)


def attention(q, k, v, amp):
    """q [Nq,H,D], k/v [Nk,H,D] -> [Nq,H,D]; flash-attn semantics in amp mode (see oracle/wan_dit.attention)."""
    from .wan_dit import attention as _att
    return _att(q, k, v, amp=amp)


def bsa_attention(q, k, v, grid_q, grid_k, bsa):
    """Attention._process_attn's block-sparse branch (attention.py:56-67): q [Nq,H,D], k/v [Nk,H,D] in (t,h,w) order."""
    from . import longcat_bsa
    t = lambda z: z.permute(1, 0, 2).unsqueeze(0)
    o = longcat_bsa.bsa_3d(t(q), t(k), t(v), grid_q, grid_k, sparsity=bsa.get("sparsity", 0.875), cdf_threshold=bsa.get("cdf_threshold"),
                           chunk_q=tuple(bsa.get("chunk_3d_shape_q", (4, 4, 8))), chunk_k=tuple(bsa.get("chunk_3d_shape_k", (4, 4, 8))))
    return o[0].permute(1, 0, 2)


def block_forward(P, cfg: LongCatConfig, i, x, y, t, grid, num_cond, amp, bsa=None, return_kv=False, kv_cache=None, skip_crs=False):
    """x [N, C] (one sample), y [M, C] valid text tokens, t [T, A] fp32.  ``bsa``: the checkpoint's bsa_params dict when
    the block-sparse self-attention of the refine pass is enabled (Attention.enable_bsa, attention.py:56), else None.
    ``return_kv``: also return (k, v) [N,H,D] after the q/k norm and BEFORE RoPE (attention.py:119-120); ``kv_cache``: such a
    pair from the clean condition frames - the video-continuation path Attention.forward_with_kv_cache (:149-181): x holds
    the noise frames only, keys/values are [cache | own] and RoPE runs over T + num_cond frames; cross-attention then sees
    every token (longcat_video_dit.py:105-108).  ``skip_crs``: no cross-attention (the caching pass, :104)."""
    b = f"blocks.{i}."
    C, Hn, D = cfg.hidden_size, cfg.num_heads, cfg.head_dim
    T = grid[0]
    N = x.shape[0]
    per = N // T
    x_dtype = x.dtype
    mod = lin_fp32_island(F.silu(t), P[b + "adaLN_modulation.1.weight"], P[b + "adaLN_modulation.1.bias"], amp)   # [T, 6C]
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = [m.unsqueeze(1) for m in mod.chunk(6, dim=-1)]                               # [T,1,C]

    def modulate(xx, shift, scale):
        n = F.layer_norm(xx.view(T, per, C).float(), (C,), None, None, 1e-6)
        return (n * (scale + 1) + shift).to(xx.dtype).view(N, C)

    # self attention
    xm = modulate(x, sh_a, sc_a)
    qkv = lin(xm, P[b + "attn.qkv.weight"], P[b + "attn.qkv.bias"], amp).view(N, 3, Hn, D)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    q, k = rms_head(q, P[b + "attn.q_norm.weight"], amp), rms_head(k, P[b + "attn.k_norm.weight"], amp)
    kv_out = (k.clone(), v.clone()) if return_kv else None
    if kv_cache is not None:
        k_c, v_c = kv_cache
        k = torch.cat([k_c.to(k.dtype), k], dim=0)
        v = torch.cat([v_c.to(v.dtype), v], dim=0)
        fr = rope_freqs(D, (T + num_cond, grid[1], grid[2]))
        q, k = rope_apply(q, fr[-N:]), rope_apply(k, fr)
        grid_k = (T + num_cond, grid[1], grid[2])
        a = bsa_attention(q, k, v, grid, grid_k, bsa) if (bsa is not None and T > 1) else attention(q, k, v, amp)
        num_cond = 0                                                  # from here on every token of x is a noise token
    else:
        fr = rope_freqs(D, grid)
        q, k = rope_apply(q, fr), rope_apply(k, fr)
    nc = num_cond * per
    if kv_cache is not None:
        pass
    elif bsa is not None and T > 1:                                   # "bsa will not be used in image training / sampling"
        gH, gW = grid[1], grid[2]
        if nc > 0:
            a = torch.cat([bsa_attention(q[:nc], k[:nc], v[:nc], (num_cond, gH, gW), (num_cond, gH, gW), bsa),
                           bsa_attention(q[nc:], k, v, (T - num_cond, gH, gW), (T, gH, gW), bsa)], dim=0)
        else:
            a = bsa_attention(q, k, v, grid, grid, bsa)
    elif nc > 0:
        a = torch.cat([attention(q[:nc], k[:nc], v[:nc], amp), attention(q[nc:], k, v, amp)], dim=0)
    else:
        a = attention(q, k, v, amp)
    xs = lin(a.reshape(N, C), P[b + "attn.proj.weight"], P[b + "attn.proj.bias"], amp)
    x = (x + (g_a * xs.view(T, per, C)).view(N, C)).to(x_dtype)

    # cross attention: noise tokens only; condition tokens receive zero (attention.py:262-273)
    if skip_crs:
        return _ffn(P, b, x, T, per, C, sh_m, sc_m, g_m, amp, x_dtype, kv_out)
    xn = F.layer_norm(x.float(), (C,), P[b + "pre_crs_attn_norm.weight"].float() if not amp else _rb(P[b + "pre_crs_attn_norm.weight"]),
                      P[b + "pre_crs_attn_norm.bias"].float() if not amp else _rb(P[b + "pre_crs_attn_norm.bias"]), 1e-6).to(x_dtype)
    qc = lin(xn[nc:], P[b + "cross_attn.q_linear.weight"], P[b + "cross_attn.q_linear.bias"], amp).view(-1, Hn, D)
    kv = lin(y, P[b + "cross_attn.kv_linear.weight"], P[b + "cross_attn.kv_linear.bias"], amp).view(-1, 2, Hn, D)
    kc, vc = kv[:, 0], kv[:, 1]
    qc, kc = rms_head(qc, P[b + "cross_attn.q_norm.weight"], amp), rms_head(kc, P[b + "cross_attn.k_norm.weight"], amp)
    oc = lin(attention(qc, kc, vc, amp).reshape(-1, C), P[b + "cross_attn.proj.weight"], P[b + "cross_attn.proj.bias"], amp)
    x = torch.cat([x[:nc], x[nc:] + oc], dim=0)

    return _ffn(P, b, x, T, per, C, sh_m, sc_m, g_m, amp, x_dtype, kv_out)


def _ffn(P, b, x, T, per, C, sh_m, sc_m, g_m, amp, x_dtype, kv_out):
    """SwiGLU feed-forward with modulation (longcat_video_dit.py:110-115)."""
    N = x.shape[0]
    n = F.layer_norm(x.view(T, per, C).float(), (C,), None, None, 1e-6)
    xm = (n * (sc_m + 1) + sh_m).to(x.dtype).view(N, C)
    h = F.silu(lin(xm, P[b + "ffn.w1.weight"], None, amp)) * lin(xm, P[b + "ffn.w3.weight"], None, amp)
    xs = lin(h, P[b + "ffn.w2.weight"], None, amp)
    x = (x + (g_m * xs.view(T, per, C)).view(N, C)).to(x_dtype)
    return x if kv_out is None else (x, kv_out)


def dit_forward(P, cfg: LongCatConfig, x, timestep, context, num_cond_latents: int = 1, amp: bool = True, bsa=None,
                return_kv: bool = False, kv_cache_dict=None, skip_crs_attn: bool = False):
    """LongCatVideoTransformer3DModel.forward for one sample (``return_kv`` / ``kv_cache_dict`` / ``skip_crs_attn``:
    longcat_video_dit.py:285-287, 339-356 - the video-continuation path; returns (out, {layer: (k, v)}) with return_kv).

    x [C_in, T, H, W]; timestep [T] (per latent frame, the condition frames at 0, pipeline_longcat_video.py:864-865);
    context [M, caption_channels] = the valid text tokens.  Returns fp32 [C_out, T, H, W]."""
    Cin, T, H, W = x.shape
    pt, ph, pw = cfg.patch
    grid = (T, H // ph, W // pw)
    N = grid[0] * grid[1] * grid[2]
    cols = x.view(Cin, grid[0], pt, grid[1], ph, grid[2], pw).permute(1, 3, 5, 0, 2, 4, 6).reshape(N, -1)
    dt = BF16 if amp else F32
    tok = lin(cols.to(dt), P["x_embedder.proj.weight"].flatten(1), P["x_embedder.proj.bias"], amp)
    ts = timestep.to(dt).float().flatten()                      # timestep.to(dtype) then .float() (:307,:313)
    emb = timestep_embedding(ts, cfg.frequency_embedding_size)
    t = lin_fp32_island(F.silu(lin_fp32_island(emb, P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"], amp)),
                        P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"], amp)          # [T, A] fp32
    y = lin(F.gelu(lin(context.to(dt), P["y_embedder.y_proj.0.weight"], P["y_embedder.y_proj.0.bias"], amp), approximate="tanh"),
            P["y_embedder.y_proj.2.weight"], P["y_embedder.y_proj.2.bias"], amp)
    kv_ret = {}
    for i in range(cfg.depth):
        r = block_forward(P, cfg, i, tok, y, t, grid, num_cond_latents, amp, bsa, return_kv=return_kv,
                          kv_cache=None if not kv_cache_dict else kv_cache_dict.get(i), skip_crs=skip_crs_attn)
        if return_kv:
            tok, kv_ret[i] = r
        else:
            tok = r
    mod = lin_fp32_island(F.silu(t), P["final_layer.adaLN_modulation.1.weight"], P["final_layer.adaLN_modulation.1.bias"], amp)
    shift, scale = [m.unsqueeze(1) for m in mod.chunk(2, dim=-1)]
    per = N // T
    n = F.layer_norm(tok.view(T, per, -1).float(), (cfg.hidden_size,), None, None, 1e-6)
    h = (n * (scale + 1) + shift).to(tok.dtype).view(N, -1)
    out = lin_fp32_island(h, P["final_layer.linear.weight"], P["final_layer.linear.bias"], amp)   # fp32 [N, pt*ph*pw*Cout]
    c = cfg.out_channels
    u = out.view(grid[0], grid[1], grid[2], pt, ph, pw, c).permute(6, 0, 3, 1, 4, 2, 5)
    out = u.reshape(c, grid[0] * pt, grid[1] * ph, grid[2] * pw).to(F32)
    return (out, kv_ret) if return_kv else out
