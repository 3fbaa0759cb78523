"""The once-per-video encoders on the engine: UMT5-XXL text encoder and CLIP ViT-H/14 image encoder.

They stand where ``pipe.text_encoder`` / ``pipe.image_encoder`` stand in the reference pipeline
(utils/pipeline_wan_i2v_clean.py:167-211 ``_get_t5_prompt_embeds``, :204-208 ``encode_image``) - transformers'
``UMT5EncoderModel`` / ``CLIPVisionModel`` there, ``wan/modules/t5.py`` / ``wan/modules/clip.py`` in the vendored Wan code;
both parameter namings are accepted.  Every Linear is the tcgen05 GEMM; norms, the T5 gated GELU and the two small-head
attentions (head_dim 64 / 80, not the 128 the DiT kernels are built for) are ``encoder_ops.cu`` / ``dit_ops.cu`` kernels.
The arithmetic is the bf16 module's: bf16 Linears with fp32 accumulation, bf16 residual stream, fp32 norm statistics, fp32
softmax - ``oracle/encoders.py`` reproduces the reference's bf16 T5Encoder bit for bit on the CPU and is what the GPU tests
compare with.  Tokenisation (sentencepiece) and the image resize / normalisation stay on the host.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict

import torch

from . import lib

BF, F32 = torch.bfloat16, torch.float32


def _t5_vendored_names(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """transformers ``UMT5EncoderModel`` names -> ``wan/modules/t5.py`` T5Encoder names (a vendored dict passes through)."""
    if not any(k.startswith(("encoder.block.", "shared.", "encoder.embed_tokens.")) for k in sd):
        return sd
    sub = {"layer.0.SelfAttention.q": "attn.q", "layer.0.SelfAttention.k": "attn.k", "layer.0.SelfAttention.v": "attn.v",
           "layer.0.SelfAttention.o": "attn.o", "layer.0.SelfAttention.relative_attention_bias": "pos_embedding.embedding",
           "layer.0.layer_norm": "norm1", "layer.1.layer_norm": "norm2", "layer.1.DenseReluDense.wi_0": "ffn.gate.0",
           "layer.1.DenseReluDense.wi_1": "ffn.fc1", "layer.1.DenseReluDense.wo": "ffn.fc2"}
    out = {}
    for k, v in sd.items():
        if k in ("shared.weight", "encoder.embed_tokens.weight"):
            out["token_embedding.weight"] = v
        elif k == "encoder.final_layer_norm.weight":
            out["norm.weight"] = v
        elif k.startswith("encoder.block."):
            _, _, i, rest = k.split(".", 3)
            for a, b in sub.items():
                if rest.startswith(a + "."):
                    out[f"blocks.{i}.{b}{rest[len(a):]}"] = v
                    break
    return out


def relative_position_bucket(L: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """Bucket of every relative distance d = j - i in [-(L-1), L-1], index d + L - 1 (T5RelativeEmbedding, t5.py:247-264,
    bidirectional)."""
    rel = torch.arange(-(L - 1), L)
    nb = num_buckets // 2
    buckets = (rel > 0).long() * nb
    a = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(a.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return (buckets + torch.where(a < max_exact, a, large)).to(torch.int32)


class WfT5Encoder:
    """``encoder(ids [B,L] int64, mask [B,L]) -> SimpleNamespace(last_hidden_state=[B,L,dim] bf16)``."""

    def __init__(self, state_dict, device, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32,
                 shared_pos=False):
        sd = _t5_vendored_names(state_dict)
        self.device, self.dtype = torch.device(device), BF
        self.dim, self.dim_attn, self.dim_ffn, self.heads, self.layers, self.buckets = dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets
        if dim_attn // num_heads != 64:
            raise lib.WfError("WfT5Encoder: head_dim must be 64 (UMT5-XXL: 4096 / 64 heads)")
        self.config = SimpleNamespace(d_model=dim)
        dev = self.device
        bf = lambda k: sd[k].to(device=dev, dtype=BF).contiguous()
        f32_of_bf = lambda k: sd[k].to(device=dev, dtype=BF).to(F32).contiguous()
        self.emb = bf("token_embedding.weight")
        self.norm = f32_of_bf("norm.weight")
        self.blocks = []
        for i in range(num_layers):
            p = f"blocks.{i}."
            b = SimpleNamespace()
            b.n1, b.n2 = f32_of_bf(p + "norm1.weight"), f32_of_bf(p + "norm2.weight")
            b.qkv = torch.cat([sd[p + "attn.q.weight"], sd[p + "attn.k.weight"], sd[p + "attn.v.weight"]], dim=0).to(device=dev, dtype=BF).contiguous()
            b.o = bf(p + "attn.o.weight")
            b.gate_fc1 = torch.cat([sd[p + "ffn.gate.0.weight"], sd[p + "ffn.fc1.weight"]], dim=0).to(device=dev, dtype=BF).contiguous()
            b.fc2 = bf(p + "ffn.fc2.weight")
            b.pos = bf("pos_embedding.embedding.weight" if shared_pos else p + "pos_embedding.embedding.weight")     # [buckets, heads]
            self.blocks.append(b)
        self._bucket = {}

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def __call__(self, input_ids, attention_mask=None, **kw):
        ids = input_ids.to(self.device)
        B, L = ids.shape
        if L > 1024:
            raise lib.WfError("WfT5Encoder: at most 1024 tokens")
        mask = torch.ones_like(ids) if attention_mask is None else attention_mask.to(self.device)
        if L not in self._bucket:
            self._bucket[L] = relative_position_bucket(L, self.buckets).to(self.device)
        bucket = self._bucket[L]
        C, A, Fd = self.dim, self.dim_attn, self.dim_ffn
        e = lambda *s: torch.empty(*s, dtype=BF, device=self.device)
        h, qkv, att, gf, ff = e(L, C), e(L, 3 * A), e(L, A), e(L, 2 * Fd), e(L, Fd)
        outs = []
        for s in range(B):
            valid = mask[s] != 0
            n_valid = int(valid.sum())
            if n_valid == 0 or not bool(valid[:n_valid].all()):
                raise lib.WfError("WfT5Encoder: the attention mask must be a non-empty prefix (right padding), as the tokenizer produces")
            x = self.emb[ids[s]].contiguous()                                   # [L, C] bf16 residual stream
            for b in self.blocks:
                h.copy_(x)
                lib.rms_norm_rope_(h, b.n1, 1e-6, None)                         # T5LayerNorm (t5.py:61-66)
                lib.gemm_bf16(h, b.qkv, None, qkv, lib.EPI_BF16)
                lib.attention_small(qkv[:, :A], qkv[:, A:2 * A], qkv[:, 2 * A:], att, self.heads, 64, mode=0, n_valid=n_valid,
                                    bias_emb=b.pos, bias_bucket=bucket)
                lib.gemm_bf16(att, b.o, None, x, lib.EPI_RESID_BF16)
                h.copy_(x)
                lib.rms_norm_rope_(h, b.n2, 1e-6, None)
                lib.gemm_bf16(h, b.gate_fc1, None, gf, lib.EPI_BF16)
                lib.geglu_bf16(gf, ff)
                lib.gemm_bf16(ff, b.fc2, None, x, lib.EPI_RESID_BF16)
            lib.rms_norm_rope_(x, self.norm, 1e-6, None)
            outs.append(x)
        return SimpleNamespace(last_hidden_state=torch.stack(outs))


def t5_prompt_embeds(encoder, input_ids, attention_mask, max_sequence_length: int = 512, dtype=BF):
    """``_get_t5_prompt_embeds`` after tokenisation (utils/pipeline_wan_i2v_clean.py:195-206): hidden states of the valid
    tokens, zero beyond them, padded to ``max_sequence_length``."""
    hs = encoder(input_ids, attention_mask).last_hidden_state.to(dtype)
    lens = attention_mask.gt(0).sum(dim=1).long()
    out = hs.new_zeros(hs.shape[0], max_sequence_length, hs.shape[2])
    for i, n in enumerate(lens.tolist()):
        out[i, :n] = hs[i, :n]
    return out


class WfCLIPVisionEncoder:
    """``encoder(pixel_values [B,3,S,S]) -> SimpleNamespace(hidden_states=(..., penultimate [B, 1+n, dim]))``: the ViT of
    wan/modules/clip.py:209-300 evaluated up to the last-but-one block (``use_31_block``; ``hidden_states[-2]`` of
    transformers' CLIPVisionModel, which the pipeline takes at :204-208)."""

    def __init__(self, state_dict, device, image_size=224, patch_size=14, dim=1280, mlp_ratio=4, num_heads=16, num_layers=32, eps=1e-5):
        sd = _clip_vendored_names(state_dict)
        self.device, self.dtype = torch.device(device), BF
        self.patch, self.dim, self.heads, self.layers, self.eps = patch_size, dim, num_heads, num_layers, eps
        if dim // num_heads != 80:
            raise lib.WfError("WfCLIPVisionEncoder: head_dim must be 80 (ViT-H/14: 1280 / 16 heads)")
        self.config = SimpleNamespace(hidden_size=dim, image_size=image_size, patch_size=patch_size)
        dev = self.device
        bf = lambda k: sd[k].to(device=dev, dtype=BF).contiguous()
        f32_of_bf = lambda k: sd[k].to(device=dev, dtype=BF).to(F32).contiguous()
        pw = sd["patch_embedding.weight"].flatten(1)
        self.k_pad = (pw.shape[1] + 7) // 8 * 8                                 # GEMM rows are 16-byte multiples (3*14*14 = 588 -> 592)
        self.patch_w = torch.nn.functional.pad(pw, (0, self.k_pad - pw.shape[1])).to(device=dev, dtype=BF).contiguous()
        self.cls = bf("cls_embedding").view(1, dim)
        self.pos = bf("pos_embedding")[0]
        self.pre_w, self.pre_b = f32_of_bf("pre_norm.weight"), f32_of_bf("pre_norm.bias")
        self.blocks = []
        for i in range(num_layers - 1):                                         # the last block is never evaluated
            p = f"transformer.{i}."
            b = SimpleNamespace(n1w=f32_of_bf(p + "norm1.weight"), n1b=f32_of_bf(p + "norm1.bias"), n2w=f32_of_bf(p + "norm2.weight"),
                                n2b=f32_of_bf(p + "norm2.bias"), qkv_w=bf(p + "attn.to_qkv.weight"), qkv_b=bf(p + "attn.to_qkv.bias"),
                                proj_w=bf(p + "attn.proj.weight"), proj_b=bf(p + "attn.proj.bias"), f0_w=bf(p + "mlp.0.weight"),
                                f0_b=bf(p + "mlp.0.bias"), f2_w=bf(p + "mlp.2.weight"), f2_b=bf(p + "mlp.2.bias"))
            self.blocks.append(b)

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def __call__(self, pixel_values, output_hidden_states=True, **kw):
        px = pixel_values.to(self.device, F32)
        B, _, S, _ = px.shape
        p, C = self.patch, self.dim
        n = (S // p) ** 2
        L = n + 1
        if L != self.pos.shape[0]:
            raise lib.WfError("WfCLIPVisionEncoder: the image size does not match the position table")
        e = lambda *s: torch.empty(*s, dtype=BF, device=self.device)
        x, h, qkv, att, ff = e(L, C), e(L, C), e(L, 3 * C), e(L, C), e(L, self.blocks[0].f0_w.shape[0])
        outs = []
        for s in range(B):
            cols = torch.nn.functional.unfold(px[s:s + 1], p, stride=p)[0].t().to(BF)                  # [n, 3*p*p] (layout only)
            cols = torch.nn.functional.pad(cols, (0, self.k_pad - cols.shape[1])).contiguous()
            lib.gemm_bf16(cols, self.patch_w, None, h[1:], lib.EPI_BF16)
            h[0].copy_(self.cls[0])
            h.add_(self.pos)                                                    # bf16 adds of the bf16 module (:291)
            lib.layer_norm(h, x, self.eps, weight=self.pre_w, bias=self.pre_b)
            for b in self.blocks:
                lib.layer_norm(x, h, self.eps, weight=b.n1w, bias=b.n1b)
                lib.gemm_bf16(h, b.qkv_w, b.qkv_b, qkv, lib.EPI_BF16)
                lib.attention_small(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], att, self.heads, 80, mode=1, scale=80 ** -0.5)
                lib.gemm_bf16(att, b.proj_w, b.proj_b, x, lib.EPI_RESID_BF16)
                lib.layer_norm(x, h, self.eps, weight=b.n2w, bias=b.n2b)
                lib.gemm_bf16(h, b.f0_w, b.f0_b, ff, lib.EPI_BF16)
                lib.gelu_erf_bf16_(ff)
                lib.gemm_bf16(ff, b.f2_w, b.f2_b, x, lib.EPI_RESID_BF16)
            outs.append(x.clone())
        pen = torch.stack(outs)
        return SimpleNamespace(hidden_states=(pen, pen, None), last_hidden_state=None)


def _clip_vendored_names(sd):
    """transformers ``CLIPVisionModel`` names -> wan/modules/clip.py VisionTransformer names (a vendored dict, optionally
    prefixed with ``visual.``, passes through with the prefix removed)."""
    if any(k.startswith("visual.") for k in sd):
        sd = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
    if not any(k.startswith("vision_model.") for k in sd):
        return sd
    out = {}
    qkv = {}
    for k, v in sd.items():
        k = k[len("vision_model."):] if k.startswith("vision_model.") else k
        if k == "embeddings.class_embedding":
            out["cls_embedding"] = v.view(1, 1, -1)
        elif k == "embeddings.patch_embedding.weight":
            out["patch_embedding.weight"] = v
        elif k == "embeddings.position_embedding.weight":
            out["pos_embedding"] = v.unsqueeze(0)
        elif k.startswith("pre_layrnorm."):
            out["pre_norm." + k.split(".", 1)[1]] = v
        elif k.startswith("encoder.layers."):
            _, _, i, rest = k.split(".", 3)
            p = f"transformer.{i}."
            m = {"layer_norm1": "norm1", "layer_norm2": "norm2", "self_attn.out_proj": "attn.proj", "mlp.fc1": "mlp.0", "mlp.fc2": "mlp.2"}
            done = False
            for a, b in m.items():
                if rest.startswith(a + "."):
                    out[p + b + rest[len(a):]] = v; done = True
                    break
            if not done and rest.startswith("self_attn."):
                _, which, kind = rest.split(".")                                # q_proj / k_proj / v_proj . weight / bias
                qkv.setdefault((i, kind), {})[which] = v
    for (i, kind), d in qkv.items():
        out[f"transformer.{i}.attn.to_qkv.{kind}"] = torch.cat([d["q_proj"], d["k_proj"], d["v_proj"]], dim=0)
    return out
