"""Oracle: Wan2.1 I2V DiT forward, restated functionally on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference's vendored PyTorch + flash-attn statement of the model,
``wan_for_worldforge/wan/modules/model.py`` (WanModel.forward :493-582,
WanAttentionBlock.forward :278-317, WanSelfAttention :130-159,
WanI2VCrossAttention :202-229, rope_apply :43-70, WanRMSNorm :81-89,
WanLayerNorm :97-102, Head :337-347, MLPProj :363-369, unpatchify :584-607)
and ``wan/modules/attention.py`` flash_attention :24-130.

Two numeric modes:

* ``amp=False``: everything in fp32 (what the reference computes when it is
  run in fp32 on the CPU).  Used to pin this restatement against the imported
  reference at ~1e-5.
* ``amp=True``: the dtype flow the reference has on the GPU, where upstream Wan
  runs fp32 master weights under ``torch.autocast(bf16)``
  (wan/image2video.py:258-330): every nn.Linear / the patch Conv3d takes bf16
  operands, accumulates in fp32 and rounds its output once to bf16; LayerNorm,
  the modulation arithmetic, the residual stream and the head stay fp32
  (the ``amp.autocast(dtype=torch.float32)`` regions, model.py:297-313,344-346);
  RMSNorm rounds its normalised value to bf16 before the fp32 weight multiply
  (:86); RoPE is evaluated in float64 and rounded to fp32 (:55-70);
  flash-attention sees bf16 q,k,v and returns bf16 (attention.py:60-79,130).
  A bf16 GEMM is modelled as exact products of the bf16 operands with fp32
  accumulation and a single rounding - the accumulation ORDER is the only
  freedom a GPU kernel has against this oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
F32 = torch.float32


@dataclass
class DitConfig:
    dim: int = 5120
    ffn_dim: int = 13824
    num_heads: int = 40
    num_layers: int = 40
    in_dim: int = 36          # 16 latent + 4 mask + 16 condition channels (model.py:531)
    out_dim: int = 16
    freq_dim: int = 256
    text_dim: int = 4096
    text_len: int = 512
    img_dim: int = 1280
    img_len: int = 257
    patch: Tuple[int, int, int] = (1, 2, 2)
    eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


WAN_I2V_14B = DitConfig()   # wan/configs/wan_i2v_14B.py:27-36


def param_shapes(cfg: DitConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameter names (vendored WanModel state-dict keys) and shapes."""
    d, f = cfg.dim, cfg.ffn_dim
    pt, ph, pw = cfg.patch
    s: Dict[str, Tuple[int, ...]] = {
        "patch_embedding.weight": (d, cfg.in_dim, pt, ph, pw),
        "patch_embedding.bias": (d,),
        "text_embedding.0.weight": (d, cfg.text_dim), "text_embedding.0.bias": (d,),
        "text_embedding.2.weight": (d, d), "text_embedding.2.bias": (d,),
        "time_embedding.0.weight": (d, cfg.freq_dim), "time_embedding.0.bias": (d,),
        "time_embedding.2.weight": (d, d), "time_embedding.2.bias": (d,),
        "time_projection.1.weight": (6 * d, d), "time_projection.1.bias": (6 * d,),
        "img_emb.proj.0.weight": (cfg.img_dim,), "img_emb.proj.0.bias": (cfg.img_dim,),
        "img_emb.proj.1.weight": (cfg.img_dim, cfg.img_dim), "img_emb.proj.1.bias": (cfg.img_dim,),
        "img_emb.proj.3.weight": (d, cfg.img_dim), "img_emb.proj.3.bias": (d,),
        "img_emb.proj.4.weight": (d,), "img_emb.proj.4.bias": (d,),
        "head.modulation": (1, 2, d),
        "head.head.weight": (cfg.out_dim * pt * ph * pw, d),
        "head.head.bias": (cfg.out_dim * pt * ph * pw,),
    }
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        s[b + "modulation"] = (1, 6, d)
        for att in ("self_attn", "cross_attn"):
            for lin in ("q", "k", "v", "o"):
                s[b + f"{att}.{lin}.weight"] = (d, d)
                s[b + f"{att}.{lin}.bias"] = (d,)
            s[b + f"{att}.norm_q.weight"] = (d,)
            s[b + f"{att}.norm_k.weight"] = (d,)
        for lin in ("k_img", "v_img"):
            s[b + f"cross_attn.{lin}.weight"] = (d, d)
            s[b + f"cross_attn.{lin}.bias"] = (d,)
        s[b + "cross_attn.norm_k_img.weight"] = (d,)
        s[b + "norm3.weight"] = (d,)
        s[b + "norm3.bias"] = (d,)
        s[b + "ffn.0.weight"] = (f, d); s[b + "ffn.0.bias"] = (f,)
        s[b + "ffn.2.weight"] = (d, f); s[b + "ffn.2.bias"] = (d,)
    return s


def init_params(cfg: DitConfig, seed: int = 1234, device="cpu") -> Dict[str, torch.Tensor]:
    """Deterministic random-init fp32 master weights.

    Matrices ~ N(0, 0.02^2); norm gains 1 + N(0, 0.05^2); norm/linear biases
    N(0, 0.02^2); modulation tables randn/sqrt(dim) (model.py:276,335).  The
    head projection is NOT zero (the reference zero-inits it, model.py:631,
    which would make every parity test vacuous).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("modulation"):
            w = torch.randn(shape, generator=g) / math.sqrt(cfg.dim)
        elif len(shape) == 1 and name.endswith("weight"):
            w = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            w = 0.02 * torch.randn(shape, generator=g)
        out[name] = w.to(device)
    return out


# --------------------------------------------------------------------------
# primitive ops with the reference's rounding points
# --------------------------------------------------------------------------

def _rb(x: torch.Tensor) -> torch.Tensor:
    """Values of x rounded to bf16, held in fp32."""
    return x.to(BF16).to(F32)


def amp_linear(x, w, b, amp: bool):
    """nn.Linear under autocast(bf16): bf16 operands, fp32 accumulate, bf16 out."""
    if not amp:
        return F.linear(x.to(F32), w, b)
    y = F.linear(_rb(x), _rb(w), None if b is None else _rb(b))
    return y.to(BF16)


def layer_norm(x, eps, w=None, b=None):
    """WanLayerNorm.forward (model.py:97-102): fp32 LN, cast back to x.dtype."""
    y = F.layer_norm(x.to(F32), (x.shape[-1],), w, b, eps)
    return y.to(x.dtype)


def rms_norm(x, w, eps):
    """WanRMSNorm.forward (model.py:81-89)."""
    xf = x.to(F32)
    y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return y.to(x.dtype) * w


def rope_table(head_dim: int, max_pos: int = 1024, theta: float = 10000.0):
    """The three complex128 frequency tables of WanModel.__init__ (model.py:479-485)
    already split the way rope_apply splits them (:47)."""
    def one(dim):
        inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim)
        ang = torch.outer(torch.arange(max_pos, dtype=torch.float64), inv)
        return torch.polar(torch.ones_like(ang), ang)
    d6 = head_dim // 6
    return one(head_dim - 4 * d6), one(2 * d6), one(2 * d6)


def rope_angles(head_dim: int, grid: Tuple[int, int, int]) -> torch.Tensor:
    """Per-token rotation as complex128 [L, head_dim/2] (model.py:57-62)."""
    f, h, w = grid
    tf, th, tw = rope_table(head_dim)
    return torch.cat([
        tf[:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        th[:h].view(1, h, 1, -1).expand(f, h, w, -1),
        tw[:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(f * h * w, -1)


def rope_apply(x, grid):
    """x [L, n, d] -> float64 complex rotation -> fp32 (model.py:43-70)."""
    L, n, d = x.shape
    xc = torch.view_as_complex(x.to(torch.float64).reshape(L, n, d // 2, 2))
    rot = rope_angles(d, grid).unsqueeze(1)
    return torch.view_as_real(xc * rot).flatten(2).to(F32)


def attention(q, k, v, amp: bool, p_bf16: bool = True):
    """Non-causal softmax(q k^T / sqrt(d)) v per head; q [Lq,n,d], k,v [Lk,n,d].

    amp=True models flash-attn on bf16 inputs (attention.py:60-130): fp32
    scores and statistics, probabilities rounded to bf16 before the PV
    product, fp32 accumulation, bf16 output (returned widened to q.dtype like
    ``x.type(out_dtype)`` does).
    """
    out_dtype = q.dtype
    if amp:
        q, k, v = _rb(q), _rb(k), _rb(v)
    else:
        q, k, v = q.to(F32), k.to(F32), v.to(F32)
    scale = q.shape[-1] ** -0.5
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    m = s.amax(dim=-1, keepdim=True)
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    if amp and p_bf16:
        p = _rb(p)
    o = torch.einsum("hqk,khd->qhd", p, v) / l.permute(1, 0, 2)
    if amp:
        o = o.to(BF16)
    return o.to(out_dtype)


def sinusoid(freq_dim: int, t: torch.Tensor) -> torch.Tensor:
    """sinusoidal_embedding_1d (model.py:18-28), float64."""
    half = freq_dim // 2
    pos = t.to(torch.float64).reshape(-1)
    ang = torch.outer(pos, torch.pow(10000.0, -torch.arange(half, dtype=torch.float64) / half))
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


# --------------------------------------------------------------------------
# the forward
# --------------------------------------------------------------------------

def time_embed(P, cfg: DitConfig, t):
    """e [1,dim], e0 [1,6,dim] fp32 (model.py:546-550)."""
    s = sinusoid(cfg.freq_dim, torch.as_tensor(t)).to(F32)
    e = F.linear(F.silu(F.linear(s, P["time_embedding.0.weight"], P["time_embedding.0.bias"])),
                 P["time_embedding.2.weight"], P["time_embedding.2.bias"])
    e0 = F.linear(F.silu(e), P["time_projection.1.weight"], P["time_projection.1.bias"])
    return e, e0.unflatten(1, (6, cfg.dim))


def embed_context(P, cfg: DitConfig, context, clip_fea, amp: bool):
    """[img_len + text_len, dim] cross-attention context (model.py:553-563)."""
    ctx = context
    if ctx.shape[0] < cfg.text_len:
        ctx = torch.cat([ctx, ctx.new_zeros(cfg.text_len - ctx.shape[0], ctx.shape[1])])
    h = amp_linear(ctx, P["text_embedding.0.weight"], P["text_embedding.0.bias"], amp)
    h = F.gelu(h, approximate="tanh")
    txt = amp_linear(h, P["text_embedding.2.weight"], P["text_embedding.2.bias"], amp)
    c = F.layer_norm(clip_fea.to(F32), (cfg.img_dim,), P["img_emb.proj.0.weight"],
                     P["img_emb.proj.0.bias"], 1e-5)
    c = amp_linear(c, P["img_emb.proj.1.weight"], P["img_emb.proj.1.bias"], amp)
    c = F.gelu(c)
    c = amp_linear(c, P["img_emb.proj.3.weight"], P["img_emb.proj.3.bias"], amp)
    img = F.layer_norm(c.to(F32), (cfg.dim,), P["img_emb.proj.4.weight"],
                       P["img_emb.proj.4.bias"], 1e-5)
    return torch.cat([img, txt.to(F32)], dim=0)


def block_forward(P, cfg: DitConfig, i: int, x, e0, grid, context, amp: bool, p_bf16=True):
    """One WanAttentionBlock (model.py:278-317); x [L, dim]."""
    b = f"blocks.{i}."
    n, d = cfg.num_heads, cfg.head_dim
    L = x.shape[0]
    e = (P[b + "modulation"] + e0)[0]          # [6, dim] fp32
    lin = lambda t_, name: amp_linear(t_, P[b + name + ".weight"], P[b + name + ".bias"], amp)

    # self attention
    h = layer_norm(x, cfg.eps).to(F32) * (1 + e[1]) + e[0]
    q = rms_norm(lin(h, "self_attn.q"), P[b + "self_attn.norm_q.weight"], cfg.eps).view(L, n, d)
    k = rms_norm(lin(h, "self_attn.k"), P[b + "self_attn.norm_k.weight"], cfg.eps).view(L, n, d)
    v = lin(h, "self_attn.v").view(L, n, d)
    a = attention(rope_apply(q, grid), rope_apply(k, grid), v, amp, p_bf16)
    y = lin(a.flatten(1), "self_attn.o")
    x = x + y * e[2]

    # cross attention (image keys first, then text keys; the two results are added)
    hq = layer_norm(x, cfg.eps, P[b + "norm3.weight"], P[b + "norm3.bias"])
    ctx_img, ctx_txt = context[:cfg.img_len], context[cfg.img_len:]
    q = rms_norm(lin(hq, "cross_attn.q"), P[b + "cross_attn.norm_q.weight"], cfg.eps).view(L, n, d)
    k = rms_norm(lin(ctx_txt, "cross_attn.k"), P[b + "cross_attn.norm_k.weight"], cfg.eps).view(-1, n, d)
    v = lin(ctx_txt, "cross_attn.v").view(-1, n, d)
    ki = rms_norm(lin(ctx_img, "cross_attn.k_img"), P[b + "cross_attn.norm_k_img.weight"], cfg.eps).view(-1, n, d)
    vi = lin(ctx_img, "cross_attn.v_img").view(-1, n, d)
    a_img = attention(q, ki, vi, amp, p_bf16)
    a_txt = attention(q, k, v, amp, p_bf16)
    x = x + lin((a_txt + a_img).flatten(1), "cross_attn.o")

    # feed forward
    h = layer_norm(x, cfg.eps).to(F32) * (1 + e[4]) + e[3]
    y = lin(F.gelu(lin(h, "ffn.0"), approximate="tanh"), "ffn.2")
    x = x + y * e[5]
    return x


def patchify(x, cfg: DitConfig):
    """[C, F, H, W] -> im2col rows [L, C*pt*ph*pw] in the Conv3d weight's
    (c, pt, ph, pw) order, tokens in (f, h, w) order (model.py:534-537)."""
    C, Fr, H, W = x.shape
    pt, ph, pw = cfg.patch
    g = (Fr // pt, H // ph, W // pw)
    u = x.view(C, g[0], pt, g[1], ph, g[2], pw).permute(1, 3, 5, 0, 2, 4, 6)
    return u.reshape(g[0] * g[1] * g[2], C * pt * ph * pw), g


def unpatchify(y, cfg: DitConfig, grid):
    """[L, pt*ph*pw*C_out] -> [C_out, F, H, W] (model.py:600-607)."""
    c = cfg.out_dim
    pt, ph, pw = cfg.patch
    u = y.view(*grid, pt, ph, pw, c)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(c, grid[0] * pt, grid[1] * ph, grid[2] * pw)


def dit_forward(P, cfg: DitConfig, x, t, context, clip_fea, amp: bool = True,
                p_bf16: bool = True, return_tokens: bool = False):
    """WanModel.forward for one sample (model.py:493-582).

    x [in_dim, F, H, W] (latents already concatenated with the condition),
    t scalar timestep, context [<=text_len, text_dim], clip_fea [img_len, img_dim].
    Returns fp32 [out_dim, F, H, W].
    """
    cols, grid = patchify(x, cfg)
    w = P["patch_embedding.weight"].flatten(1)
    tok = amp_linear(cols, w, P["patch_embedding.bias"], amp)      # bf16 in amp mode
    e, e0 = time_embed(P, cfg, t)
    ctx = embed_context(P, cfg, context, clip_fea, amp)
    for i in range(cfg.num_layers):
        tok = block_forward(P, cfg, i, tok, e0, grid, ctx, amp, p_bf16)
    eh = (P["head.modulation"] + e.unsqueeze(1))[0]                 # [2, dim]
    h = layer_norm(tok, cfg.eps).to(F32) * (1 + eh[1]) + eh[0]
    y = F.linear(h, P["head.head.weight"], P["head.head.bias"])     # fp32 (model.py:344-346)
    if return_tokens:
        return y
    return unpatchify(y, cfg, grid).to(F32)
