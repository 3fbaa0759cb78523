"""Oracle: the once-per-video encoders, restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* UMT5-XXL text encoder: ``wan_for_worldforge/wan/modules/t5.py`` - T5LayerNorm :53-66, T5Attention :69-120 (no softmax
  scale, additive relative-position bias, masked keys at finfo.min), T5FeedForward :123-141 (fc1(x) * GELU_tanh(gate(x))),
  T5SelfAttention :144-175 (pre-norm residual block, per-block position table when shared_pos is False),
  T5RelativeEmbedding :221-264, T5Encoder.forward :303-312.  The diffusers pipeline of the entry script runs transformers'
  ``UMT5EncoderModel`` (utils/pipeline_wan_i2v_clean.py:167-211) - the same network under other parameter names.
* CLIP ViT-H/14 image encoder: ``wan/modules/clip.py`` - VisionTransformer.forward :279-300 with ``use_31_block`` (the
  penultimate hidden state, what the pipeline takes as ``hidden_states[-2]``, pipeline :204-208), AttentionBlock :146-153,
  SelfAttention :74-91, LayerNorm in fp32 :47-50.

``amp=False`` is plain fp32 (pinned against the reference classes at 1e-5); ``amp=True`` places bf16 roundings where the
bf16 modules have them (bf16 Linears with fp32 accumulation, bf16 residual stream).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

BF, F32 = torch.bfloat16, torch.float32


def _lin(x, w, b, amp):
    if not amp:
        return F.linear(x.to(F32), w, b)
    return F.linear(x.to(BF).to(F32), w.to(BF).to(F32), None if b is None else b.to(BF).to(F32)).to(BF)


# ------------------------------------------------------------------------------------------------------------ UMT5
def relative_position_bucket(rel_pos: torch.Tensor, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """T5RelativeEmbedding._relative_position_bucket, bidirectional (:247-264)."""
    nb = num_buckets // 2
    rel_buckets = (rel_pos > 0).long() * nb
    rel_pos = torch.abs(rel_pos)
    max_exact = nb // 2
    large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return rel_buckets + torch.where(rel_pos < max_exact, rel_pos, large)


def t5_layer_norm(x, w, amp, eps=1e-6):
    y = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if amp:
        return w.to(BF) * y.to(BF)
    return w * y


def gelu_tanh_expr(x):
    """GELU of t5.py:48-50 as a tensor expression (on a bf16 tensor every sub-expression rounds to bf16)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def t5_encoder(P: Dict[str, torch.Tensor], ids: torch.Tensor, mask: torch.Tensor, num_heads: int, num_layers: int,
               num_buckets: int = 32, shared_pos: bool = False, amp: bool = True) -> torch.Tensor:
    """ids, mask [L] (one sample) -> [L, dim]."""
    dt = BF if amp else F32
    x = P["token_embedding.weight"][ids].to(dt)
    L = x.shape[0]
    rel = torch.arange(L).unsqueeze(0) - torch.arange(L).unsqueeze(1)
    bucket = relative_position_bucket(rel, num_buckets)
    for i in range(num_layers):
        p = f"blocks.{i}."
        emb = P["pos_embedding.embedding.weight" if shared_pos else p + "pos_embedding.embedding.weight"].to(dt)
        bias = emb[bucket].permute(2, 0, 1)                                      # [heads, L, L]
        h = t5_layer_norm(x, P[p + "norm1.weight"], amp)
        q = _lin(h, P[p + "attn.q.weight"], None, amp).view(L, num_heads, -1)
        k = _lin(h, P[p + "attn.k.weight"], None, amp).view(L, num_heads, -1)
        v = _lin(h, P[p + "attn.v.weight"], None, amp).view(L, num_heads, -1)
        if amp:                                                                  # bf16 einsum: fp32 accumulation, bf16 result
            att = torch.einsum("inc,jnc->nij", q.float(), k.float()).to(BF)
        else:
            att = torch.einsum("inc,jnc->nij", q, k)
        att = att + bias
        att = att.masked_fill(mask.view(1, 1, -1) == 0, torch.finfo(dt).min)
        att = F.softmax(att.float(), dim=-1).to(dt)
        o = torch.einsum("nij,jnc->inc", att.float(), v.float()).to(dt).reshape(L, -1)
        x = x + _lin(o, P[p + "attn.o.weight"], None, amp)
        h = t5_layer_norm(x, P[p + "norm2.weight"], amp)
        g = gelu_tanh_expr(_lin(h, P[p + "ffn.gate.0.weight"], None, amp))
        x = x + _lin(_lin(h, P[p + "ffn.fc1.weight"], None, amp) * g, P[p + "ffn.fc2.weight"], None, amp)
    return t5_layer_norm(x, P["norm.weight"], amp)


def t5_shapes(vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets=32, shared_pos=False):
    s = {"token_embedding.weight": (vocab, dim), "norm.weight": (dim,)}
    if shared_pos:
        s["pos_embedding.embedding.weight"] = (num_buckets, num_heads)
    for i in range(num_layers):
        p = f"blocks.{i}."
        s.update({p + "norm1.weight": (dim,), p + "norm2.weight": (dim,), p + "attn.q.weight": (dim_attn, dim),
                  p + "attn.k.weight": (dim_attn, dim), p + "attn.v.weight": (dim_attn, dim), p + "attn.o.weight": (dim, dim_attn),
                  p + "ffn.gate.0.weight": (dim_ffn, dim), p + "ffn.fc1.weight": (dim_ffn, dim), p + "ffn.fc2.weight": (dim, dim_ffn)})
        if not shared_pos:
            s[p + "pos_embedding.embedding.weight"] = (num_buckets, num_heads)
    return s


def init_params(shapes, seed: int):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if len(shp) == 1 and (k.endswith("norm.weight") or "norm" in k.split(".")[-2]) and k.endswith("weight"):
            out[k] = 1.0 + 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            out[k] = 0.02 * torch.randn(shp, generator=g)
        elif "embedding" in k:
            out[k] = 0.5 * torch.randn(shp, generator=g)
        else:
            out[k] = torch.randn(shp, generator=g) / math.sqrt(shp[-1])
    return out


# ------------------------------------------------------------------------------------------------------------ CLIP ViT
def clip_shapes(image_size, patch, dim, mlp_ratio, num_layers):
    n = (image_size // patch) ** 2
    s = {"patch_embedding.weight": (dim, 3, patch, patch), "cls_embedding": (1, 1, dim), "pos_embedding": (1, n + 1, dim),
         "pre_norm.weight": (dim,), "pre_norm.bias": (dim,)}
    for i in range(num_layers):
        p = f"transformer.{i}."
        s.update({p + "norm1.weight": (dim,), p + "norm1.bias": (dim,), p + "norm2.weight": (dim,), p + "norm2.bias": (dim,),
                  p + "attn.to_qkv.weight": (3 * dim, dim), p + "attn.to_qkv.bias": (3 * dim,), p + "attn.proj.weight": (dim, dim),
                  p + "attn.proj.bias": (dim,), p + "mlp.0.weight": (int(dim * mlp_ratio), dim), p + "mlp.0.bias": (int(dim * mlp_ratio),),
                  p + "mlp.2.weight": (dim, int(dim * mlp_ratio)), p + "mlp.2.bias": (dim,)})
    return s


def clip_visual(P, img: torch.Tensor, patch: int, num_heads: int, num_layers: int, amp: bool = True, eps: float = 1e-5,
                use_31_block: bool = True) -> torch.Tensor:
    """img [3, S, S] (already resized and normalised) -> [1 + (S/patch)^2, dim] after num_layers - 1 blocks (:279-300)."""
    dt = BF if amp else F32
    dim = P["patch_embedding.weight"].shape[0]
    cols = F.unfold(img.unsqueeze(0).to(F32), patch, stride=patch)[0].t()                    # [n, 3*p*p]
    x = _lin(cols.to(dt), P["patch_embedding.weight"].flatten(1), None, amp)
    x = torch.cat([P["cls_embedding"].to(dt).view(1, dim), x.to(dt)], dim=0)
    x = x + P["pos_embedding"].to(dt)[0]
    ln = lambda t, w, b: F.layer_norm(t.float(), (dim,), w.float() if not amp else w.to(BF).float(), b.float() if not amp else b.to(BF).float(), eps).to(dt)
    x = ln(x, P["pre_norm.weight"], P["pre_norm.bias"])
    L = x.shape[0]
    for i in range(num_layers - 1 if use_31_block else num_layers):
        p = f"transformer.{i}."
        h = ln(x, P[p + "norm1.weight"], P[p + "norm1.bias"])
        qkv = _lin(h, P[p + "attn.to_qkv.weight"], P[p + "attn.to_qkv.bias"], amp).view(L, 3, num_heads, -1)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        s = torch.einsum("inc,jnc->nij", q.float(), k.float()) * (q.shape[-1] ** -0.5)
        pr = F.softmax(s, dim=-1)
        if amp:
            pr = pr.to(BF).float()
        o = torch.einsum("nij,jnc->inc", pr, v.float()).to(dt).reshape(L, dim)
        x = x + _lin(o, P[p + "attn.proj.weight"], P[p + "attn.proj.bias"], amp)
        h = ln(x, P[p + "norm2.weight"], P[p + "norm2.bias"])
        m = F.gelu(_lin(h, P[p + "mlp.0.weight"], P[p + "mlp.0.bias"], amp).float()).to(dt)
        x = x + _lin(m, P[p + "mlp.2.weight"], P[p + "mlp.2.bias"], amp)
    return x
