"""Wan2.1 I2V DiT forward on the sm_100a kernels: the object that stands in for ``pipe.transformer``.

Call surface of the reference pipeline (utils/pipeline_wan_i2v_clean.py:593-610):
``transformer(hidden_states[1,36,f,h,w] bf16, timestep[1] int64, encoder_hidden_states[1,512,4096],
encoder_hidden_states_image[1,257,1280], attention_kwargs=None, return_dict=False) -> (Tensor[1,16,f,h,w],)``
plus ``.dtype`` (:505) and ``.config.patch_size`` (infer_worldforge.py:219).

The arithmetic is the reference's PyTorch + flash-attn statement of the model, vendored
``WanModel`` (wan/modules/model.py:493-582) run the way upstream Wan runs it - fp32 master
weights under autocast(bf16) - with each rounding point kept (see oracle/wan_dit.py).  Weights
are accepted under both naming schemes (vendored ``blocks.N.self_attn.q`` ... and diffusers
``blocks.N.attn1.to_q`` ...).

One forward is ~18 kernel launches per block and nothing else: the host code below only
sequences C-ABI calls on the current CUDA stream over buffers it allocated once.

HBM layout (L tokens, D = dim):
  x      fp32 [L, D]      residual stream, updated in place by the GEMM epilogues
  h      bf16 [L, D]      LayerNorm+modulate output = A operand of the next GEMM
  qkv    bf16 [L, 3D]     fused q|k|v projection; RMSNorm+RoPE in place; attention reads the
                          three column slices as [L, heads, 128] without any re-packing
  att    bf16 [L, D]      attention output = A operand of the o-projection
  ff     bf16 [L, F]      GELU(ffn.0) output
  weights bf16 [N, K] (nn.Linear layout = K-major UMMA B operand), q|k|v and k|v stacked on N.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import lib
from .ctx_cache import ContextCache

BF, F32 = torch.bfloat16, torch.float32


@dataclass
class WanDitConfig:
    dim: int = 5120
    ffn_dim: int = 13824
    num_heads: int = 40
    num_layers: int = 40
    in_dim: int = 36
    out_dim: int = 16
    freq_dim: int = 256
    text_dim: int = 4096
    text_len: int = 512
    img_dim: int = 1280
    img_len: int = 257
    patch: Tuple[int, int, int] = (1, 2, 2)
    eps: float = 1e-6


WAN_I2V_14B = WanDitConfig()     # reference wan/configs/wan_i2v_14B.py:27-36

# diffusers WanTransformer3DModel -> vendored WanModel parameter names
_DIFFUSERS_TOP = {
    "condition_embedder.time_embedder.linear_1": "time_embedding.0",
    "condition_embedder.time_embedder.linear_2": "time_embedding.2",
    "condition_embedder.time_proj": "time_projection.1",
    "condition_embedder.text_embedder.linear_1": "text_embedding.0",
    "condition_embedder.text_embedder.linear_2": "text_embedding.2",
    "condition_embedder.image_embedder.norm1": "img_emb.proj.0",
    "condition_embedder.image_embedder.ff.net.0.proj": "img_emb.proj.1",
    "condition_embedder.image_embedder.ff.net.2": "img_emb.proj.3",
    "condition_embedder.image_embedder.norm2": "img_emb.proj.4",
    "proj_out": "head.head",
}
_DIFFUSERS_BLOCK = {
    "attn1.to_q": "self_attn.q", "attn1.to_k": "self_attn.k", "attn1.to_v": "self_attn.v", "attn1.to_out.0": "self_attn.o",
    "attn1.norm_q": "self_attn.norm_q", "attn1.norm_k": "self_attn.norm_k",
    "attn2.to_q": "cross_attn.q", "attn2.to_k": "cross_attn.k", "attn2.to_v": "cross_attn.v", "attn2.to_out.0": "cross_attn.o",
    "attn2.norm_q": "cross_attn.norm_q", "attn2.norm_k": "cross_attn.norm_k",
    "attn2.add_k_proj": "cross_attn.k_img", "attn2.add_v_proj": "cross_attn.v_img", "attn2.norm_added_k": "cross_attn.norm_k_img",
    "norm2": "norm3", "ffn.net.0.proj": "ffn.0", "ffn.net.2": "ffn.2",
}


def to_vendored_names(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accept a diffusers-style state dict and return it under the vendored WanModel key names."""
    if any(k.startswith("condition_embedder.") or ".attn1." in k for k in sd):
        out = {}
        for k, v in sd.items():
            if k == "scale_shift_table":
                out["head.modulation"] = v; continue
            parts = k.split(".")
            if parts[0] == "blocks":
                rest = ".".join(parts[2:])
                if rest == "scale_shift_table":
                    out[f"blocks.{parts[1]}.modulation"] = v; continue
                for src, dst in _DIFFUSERS_BLOCK.items():
                    if rest.startswith(src + "."):
                        out[f"blocks.{parts[1]}.{dst}{rest[len(src):]}"] = v; break
                else:
                    out[k] = v
                continue
            for src, dst in _DIFFUSERS_TOP.items():
                if k.startswith(src + "."):
                    out[dst + k[len(src):]] = v; break
            else:
                out[k] = v
        return out
    return sd


def rope_table(grid: Tuple[int, int, int], head_dim: int = 128, theta: float = 10000.0) -> torch.Tensor:
    """float64 [L, head_dim/2, 2] (cos, sin) of every token's rotation.

    Frequencies as WanModel builds them (model.py:32-39, 479-485): three tables over
    (head_dim - 4*(head_dim//6), 2*(head_dim//6), 2*(head_dim//6)) = (44, 42, 42) of the 128
    channels, indexed by the token's (frame, row, column) (rope_apply :57-62)."""
    f, h, w = grid
    d6 = head_dim // 6
    def ang(n, dim):
        inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim)
        return torch.outer(torch.arange(n, dtype=torch.float64), inv)
    af, ah, aw = ang(f, head_dim - 4 * d6), ang(h, 2 * d6), ang(w, 2 * d6)
    a = torch.cat([af.view(f, 1, 1, -1).expand(f, h, w, -1), ah.view(1, h, 1, -1).expand(f, h, w, -1),
                   aw.view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)
    return torch.stack([torch.cos(a), torch.sin(a)], dim=-1).contiguous()


class _Block:
    __slots__ = ("qkv_w", "qkv_b", "o_w", "o_b", "norm_q", "norm_k", "cq_w", "cq_b", "ckv_w", "ckv_b", "ckvi_w", "ckvi_b",
                 "co_w", "co_b", "cnorm_q", "cnorm_k", "cnorm_ki", "n3_w", "n3_b", "f0_w", "f0_b", "f2_w", "f2_b")


class WfWanTransformer:
    """Wan2.1-I2V DiT with bf16 tensor-core weights resident on one GPU."""

    def __init__(self, cfg: WanDitConfig, device):
        assert cfg.dim % cfg.num_heads == 0 and cfg.dim // cfg.num_heads == 128, "kernels are built for head_dim 128"
        assert cfg.patch == (1, 2, 2)
        self.cfg, self.device = cfg, torch.device(device)
        self.dtype = BF
        self.config = SimpleNamespace(patch_size=cfg.patch, in_channels=cfg.in_dim, out_channels=cfg.out_dim)
        self.blocks = []
        self._buf = {}
        self._rope = {}
        self._ctx_cache = ContextCache(4)
        self.cache_context = True
        self.calls = 0
        self.sp = None
        self._psp = {}

    def to(self, *a, **k):          # survives pipe.to("cuda")
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------------------ weights
    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], cfg: WanDitConfig, device) -> "WfWanTransformer":
        """sd: fp32 (or bf16) tensors under vendored or diffusers names; matrices are stored as bf16
        (what autocast feeds the tensor cores), norms / modulation / time MLPs / head stay fp32."""
        sd = to_vendored_names(sd)
        self = cls(cfg, device)
        dev = self.device
        mat = lambda k: sd[k].to(device=dev, dtype=BF).contiguous()
        vec32 = lambda k: sd[k].to(device=dev, dtype=F32).contiguous()
        cat_m = lambda *ks: torch.cat([sd[k].to(device=dev, dtype=BF) for k in ks], dim=0).contiguous()
        self.patch_w = sd["patch_embedding.weight"].flatten(1).to(device=dev, dtype=BF).contiguous()
        self.patch_b = mat("patch_embedding.bias")
        self.txt0_w, self.txt0_b = mat("text_embedding.0.weight"), mat("text_embedding.0.bias")
        self.txt2_w, self.txt2_b = mat("text_embedding.2.weight"), mat("text_embedding.2.bias")
        self.img_n0_w, self.img_n0_b = vec32("img_emb.proj.0.weight"), vec32("img_emb.proj.0.bias")
        self.img1_w, self.img1_b = mat("img_emb.proj.1.weight"), mat("img_emb.proj.1.bias")
        self.img3_w, self.img3_b = mat("img_emb.proj.3.weight"), mat("img_emb.proj.3.bias")
        self.img_n4_w, self.img_n4_b = vec32("img_emb.proj.4.weight"), vec32("img_emb.proj.4.bias")
        self.t0_w, self.t0_b = vec32("time_embedding.0.weight"), vec32("time_embedding.0.bias")
        self.t2_w, self.t2_b = vec32("time_embedding.2.weight"), vec32("time_embedding.2.bias")
        self.tp_w, self.tp_b = vec32("time_projection.1.weight"), vec32("time_projection.1.bias")
        self.head_mod = vec32("head.modulation").view(2, cfg.dim)
        self.head_w, self.head_b = vec32("head.head.weight"), vec32("head.head.bias")
        mods = []
        for i in range(cfg.num_layers):
            p = f"blocks.{i}."
            b = _Block()
            b.qkv_w = cat_m(p + "self_attn.q.weight", p + "self_attn.k.weight", p + "self_attn.v.weight")
            b.qkv_b = cat_m(p + "self_attn.q.bias", p + "self_attn.k.bias", p + "self_attn.v.bias")
            b.o_w, b.o_b = mat(p + "self_attn.o.weight"), mat(p + "self_attn.o.bias")
            b.norm_q, b.norm_k = vec32(p + "self_attn.norm_q.weight"), vec32(p + "self_attn.norm_k.weight")
            b.cq_w, b.cq_b = mat(p + "cross_attn.q.weight"), mat(p + "cross_attn.q.bias")
            b.ckv_w = cat_m(p + "cross_attn.k.weight", p + "cross_attn.v.weight")
            b.ckv_b = cat_m(p + "cross_attn.k.bias", p + "cross_attn.v.bias")
            b.ckvi_w = cat_m(p + "cross_attn.k_img.weight", p + "cross_attn.v_img.weight")
            b.ckvi_b = cat_m(p + "cross_attn.k_img.bias", p + "cross_attn.v_img.bias")
            b.co_w, b.co_b = mat(p + "cross_attn.o.weight"), mat(p + "cross_attn.o.bias")
            b.cnorm_q, b.cnorm_k = vec32(p + "cross_attn.norm_q.weight"), vec32(p + "cross_attn.norm_k.weight")
            b.cnorm_ki = vec32(p + "cross_attn.norm_k_img.weight")
            b.n3_w, b.n3_b = vec32(p + "norm3.weight"), vec32(p + "norm3.bias")
            b.f0_w, b.f0_b = mat(p + "ffn.0.weight"), mat(p + "ffn.0.bias")
            b.f2_w, b.f2_b = mat(p + "ffn.2.weight"), mat(p + "ffn.2.bias")
            self.blocks.append(b)
            mods.append(vec32(p + "modulation").view(6, cfg.dim))
        self.mod_all = torch.stack(mods).contiguous()             # [layers, 6, dim]
        return self

    @classmethod
    def random_init(cls, cfg: WanDitConfig, device, seed: int = 1234) -> "WfWanTransformer":
        """Random-init weights generated directly on the device (the 14B model is 33 GB in bf16;
        SURVEY.md §8d: N(0, 0.02^2) matrices, gains ~1, modulation randn/sqrt(dim), non-zero head)."""
        self = cls(cfg, device)
        g = torch.Generator(device=self.device).manual_seed(seed)
        d, f = cfg.dim, cfg.ffn_dim
        def m(n, k, dt=BF):
            return (torch.randn(n, k, generator=g, device=self.device, dtype=F32) * 0.02).to(dt)
        def v(n, dt=BF, s=0.02, base=0.0):
            return (base + torch.randn(n, generator=g, device=self.device, dtype=F32) * s).to(dt)
        self.patch_w, self.patch_b = m(d, cfg.in_dim * 4), v(d)
        self.txt0_w, self.txt0_b, self.txt2_w, self.txt2_b = m(d, cfg.text_dim), v(d), m(d, d), v(d)
        self.img_n0_w, self.img_n0_b = v(cfg.img_dim, F32, 0.05, 1.0), v(cfg.img_dim, F32)
        self.img1_w, self.img1_b = m(cfg.img_dim, cfg.img_dim), v(cfg.img_dim)
        self.img3_w, self.img3_b = m(d, cfg.img_dim), v(d)
        self.img_n4_w, self.img_n4_b = v(d, F32, 0.05, 1.0), v(d, F32)
        self.t0_w, self.t0_b, self.t2_w, self.t2_b = m(d, cfg.freq_dim, F32), v(d, F32), m(d, d, F32), v(d, F32)
        self.tp_w, self.tp_b = m(6 * d, d, F32), v(6 * d, F32)
        self.head_mod = v(2 * d, F32, 1.0 / math.sqrt(d)).view(2, d)
        self.head_w, self.head_b = m(cfg.out_dim * 4, d, F32), v(cfg.out_dim * 4, F32)
        mods = []
        for _ in range(cfg.num_layers):
            b = _Block()
            b.qkv_w, b.qkv_b, b.o_w, b.o_b = m(3 * d, d), v(3 * d), m(d, d), v(d)
            b.norm_q, b.norm_k = v(d, F32, 0.05, 1.0), v(d, F32, 0.05, 1.0)
            b.cq_w, b.cq_b, b.ckv_w, b.ckv_b = m(d, d), v(d), m(2 * d, d), v(2 * d)
            b.ckvi_w, b.ckvi_b, b.co_w, b.co_b = m(2 * d, d), v(2 * d), m(d, d), v(d)
            b.cnorm_q, b.cnorm_k, b.cnorm_ki = v(d, F32, 0.05, 1.0), v(d, F32, 0.05, 1.0), v(d, F32, 0.05, 1.0)
            b.n3_w, b.n3_b = v(d, F32, 0.05, 1.0), v(d, F32)
            b.f0_w, b.f0_b, b.f2_w, b.f2_b = m(f, d), v(f), m(d, f), v(d)
            self.blocks.append(b)
            mods.append(v(6 * d, F32, 1.0 / math.sqrt(d)).view(6, d))
        self.mod_all = torch.stack(mods).contiguous()
        return self

    # ------------------------------------------------------------------------------ buffers
    def _buffers(self, L: int, Ll: int):
        """L: tokens of the whole clip; Ll: tokens this rank owns (L / sequence-parallel world size)."""
        if (L, Ll) not in self._buf:
            c, dev = self.cfg, self.device
            e = lambda *s, dt=BF: torch.empty(*s, dtype=dt, device=dev)
            self._buf[(L, Ll)] = SimpleNamespace(
                cols=e(L, c.in_dim * 4), x=e(Ll, c.dim, dt=F32), h=e(Ll, c.dim), qkv=e(Ll, 3 * c.dim), att=e(Ll, c.dim),
                cq=e(Ll, c.dim), ca_img=e(Ll, c.dim), ff=e(Ll, c.ffn_dim),
                sinus=e(c.freq_dim, dt=F32), e1=e(c.dim, dt=F32), e=e(c.dim, dt=F32), e0=e(6 * c.dim, dt=F32),
                mod=e(c.num_layers, 6, c.dim, dt=F32), hmod=e(2, c.dim, dt=F32))
        return self._buf[(L, Ll)]

    def _context_kv(self, ctx_txt_in: torch.Tensor, ctx_img_in: torch.Tensor):
        """Per-block cross-attention K|V of the text and image context.  They depend only on the
        prompt / image embeddings and the weights, not on the latents or the timestep, so they are
        computed once per distinct embedding content and reused by every forward of the run."""
        if self.cache_context:
            kv = self._ctx_cache.get((ctx_txt_in, ctx_img_in))     # entries hold their source tensors: see ctx_cache.py
            if kv is not None:
                return kv
        c, dev = self.cfg, self.device
        e = lambda *s, dt=BF: torch.empty(*s, dtype=dt, device=dev)
        txt = ctx_txt_in.to(BF)
        if txt.shape[0] < c.text_len:
            txt = torch.cat([txt, txt.new_zeros(c.text_len - txt.shape[0], txt.shape[1])])
        t1 = lib.gemm_bf16(txt.contiguous(), self.txt0_w, self.txt0_b, e(c.text_len, c.dim), lib.EPI_GELU_BF16)
        ctx_txt = lib.gemm_bf16(t1, self.txt2_w, self.txt2_b, e(c.text_len, c.dim), lib.EPI_BF16)
        img = ctx_img_in.to(BF).contiguous()
        n0 = lib.layer_norm(img, e(c.img_len, c.img_dim), 1e-5, weight=self.img_n0_w, bias=self.img_n0_b)
        i1 = lib.gemm_bf16(n0, self.img1_w, self.img1_b, e(c.img_len, c.img_dim), lib.EPI_BF16)
        lib.gelu_erf_bf16_(i1)
        i3 = lib.gemm_bf16(i1, self.img3_w, self.img3_b, e(c.img_len, c.dim), lib.EPI_BF16)
        ctx_img = lib.layer_norm(i3, e(c.img_len, c.dim), 1e-5, weight=self.img_n4_w, bias=self.img_n4_b)
        kv = []
        for b in self.blocks:
            kt = lib.gemm_bf16(ctx_txt, b.ckv_w, b.ckv_b, e(c.text_len, 2 * c.dim), lib.EPI_BF16)
            lib.rms_norm_rope_(kt[:, :c.dim], b.cnorm_k, c.eps, None)
            ki = lib.gemm_bf16(ctx_img, b.ckvi_w, b.ckvi_b, e(c.img_len, 2 * c.dim), lib.EPI_BF16)
            lib.rms_norm_rope_(ki[:, :c.dim], b.cnorm_ki, c.eps, None)
            kv.append((kt, ki))
        if self.cache_context:
            self._ctx_cache.put((ctx_txt_in, ctx_img_in), kv)
        return kv

    # ------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None,
                 attention_kwargs=None, return_dict: bool = False):
        c = self.cfg
        if hidden_states.shape[0] != 1:
            raise NotImplementedError("batch size 1 (the reference runs cond / uncond as two forwards)")
        if not hidden_states.is_cuda:
            raise lib.WfError("WfWanTransformer runs on CUDA tensors only (no CPU fallback)")
        self.calls += 1
        hs = hidden_states[0].to(BF).contiguous()
        _, Fr, H, W = hs.shape
        grid = (Fr, H // 2, W // 2)
        L = grid[0] * grid[1] * grid[2]
        D, nh = c.dim, c.num_heads
        sp = self.sp                                   # sequence parallelism (worldforge_b200.ulysses), None on one GPU
        P, rk = (sp.world, sp.rank) if sp is not None else (1, 0)
        if L % P or nh % P:
            raise ValueError(f"{L} tokens / {nh} heads do not split over {P} ranks")
        Ll = L // P                                    # this rank owns the contiguous token range [rk*Ll, (rk+1)*Ll)
        B = self._buffers(L, Ll)
        if (grid, P, rk) not in self._rope:
            self._rope[(grid, P, rk)] = rope_table(grid)[rk * Ll:(rk + 1) * Ll].contiguous().to(self.device)
        rope = self._rope[(grid, P, rk)]
        psp = None
        if sp is not None and getattr(sp, "peer", False):
            if (L, Ll) not in self._psp:
                from . import ulysses
                for old in self._psp.values():         # another clip shape: release the previous exchange's buffers and mappings
                    old.close()
                self._psp.clear()
                try:
                    self._psp[(L, Ll)] = ulysses.PeerSequenceParallel(sp.group, L, Ll, nh, self.device)
                except ulysses.PeerSetupError as ex:   # raised on every rank alike: all of them keep the NCCL all-to-all form
                    import sys
                    print(f"[worldforge_b200] {ex}; using the NCCL all-to-all exchange", file=sys.stderr, flush=True)
                    sp.peer = False
            psp = self._psp.get((L, Ll))

        # patch embedding (bf16 token stream, held in fp32 storage)
        lib.patchify(hs, B.cols)
        lib.gemm_bf16(B.cols[rk * Ll:(rk + 1) * Ll], self.patch_w, self.patch_b, B.x, lib.EPI_F32_OF_BF16)
        # time embedding -> per-block modulation tables
        lib.time_sinusoid(timestep.reshape(-1)[:1].to(torch.int64).contiguous(), B.sinus)
        lib.gemv_f32(self.t0_w, B.sinus, self.t0_b, B.e1, silu_out=True)
        lib.gemv_f32(self.t2_w, B.e1, self.t2_b, B.e)
        lib.gemv_f32(self.tp_w, B.e, self.tp_b, B.e0, silu_in=True)
        lib.add_bcast_f32(self.mod_all, B.e0, B.mod)
        lib.add_bcast_f32(self.head_mod, B.e, B.hmod)
        kv_ctx = self._context_kv(encoder_hidden_states[0], encoder_hidden_states_image[0])

        for i, b in enumerate(self.blocks):
            e = B.mod[i]
            # self attention
            lib.layer_norm(B.x, B.h, c.eps, scale=e[1], shift=e[0], round_norm_bf16=(i == 0))
            lib.gemm_bf16(B.h, b.qkv_w, b.qkv_b, B.qkv, lib.EPI_BF16)
            if psp is not None:                        # Ulysses through NVLink peer memory: norm+RoPE+scatter, attention -> peers
                att = psp.exchange_attention(B.qkv, b.norm_q, b.norm_k, rope, c.eps)
            else:
                lib.rms_norm_rope_(B.qkv[:, :D], b.norm_q, c.eps, rope)
                lib.rms_norm_rope_(B.qkv[:, D:2 * D], b.norm_k, c.eps, rope)
                if sp is None:
                    lib.attention_bf16(B.qkv[:, :D], B.qkv[:, D:2 * D], B.qkv[:, 2 * D:], B.att, nh)
                    att = B.att
                else:
                    att = sp.attention(B.qkv, nh)      # Ulysses: heads <-> tokens all-to-all (NCCL) around the same kernel
            lib.gemm_bf16(att, b.o_w, b.o_b, B.x, lib.EPI_RESID_F32, gate=e[2])
            # cross attention: image keys, then text keys with the image result added
            lib.layer_norm(B.x, B.h, c.eps, weight=b.n3_w, bias=b.n3_b)
            lib.gemm_bf16(B.h, b.cq_w, b.cq_b, B.cq, lib.EPI_BF16)
            lib.rms_norm_rope_(B.cq, b.cnorm_q, c.eps, None)
            kt, ki = kv_ctx[i]
            lib.attention_bf16(B.cq, ki[:, :D], ki[:, D:], B.ca_img, nh)
            lib.attention_bf16(B.cq, kt[:, :D], kt[:, D:], B.att, nh, add_in=B.ca_img)
            lib.gemm_bf16(B.att, b.co_w, b.co_b, B.x, lib.EPI_RESID_F32)
            # feed forward
            lib.layer_norm(B.x, B.h, c.eps, scale=e[4], shift=e[3])
            lib.gemm_bf16(B.h, b.f0_w, b.f0_b, B.ff, lib.EPI_GELU_BF16)
            lib.gemm_bf16(B.ff, b.f2_w, b.f2_b, B.x, lib.EPI_RESID_F32, gate=e[5])

        if sp is None:
            out = torch.empty(c.out_dim, Fr, H, W, dtype=F32, device=self.device)
            lib.dit_head(B.x, B.hmod[1], B.hmod[0], self.head_w, self.head_b, out, grid, c.eps)
        else:                                          # every rank scatters its tokens into a zero canvas; sum = gather
            out = torch.zeros(c.out_dim, Fr, H, W, dtype=F32, device=self.device)
            lib.dit_head(B.x, B.hmod[1], B.hmod[0], self.head_w, self.head_b, out, grid, c.eps, tok_offset=rk * Ll)
            sp.all_reduce(out)
        return (out.unsqueeze(0).to(self.dtype),)

    def flops_per_forward(self, L: int) -> float:
        """Algorithmic FLOPs of one forward, 2 per MAC."""
        c = self.cfg
        ctx = c.text_len + c.img_len
        per = 2 * L * (6 * c.dim ** 2 + 2 * c.dim * c.ffn_dim) + 4 * c.dim ** 2 * ctx + 4 * L * L * c.dim + 4 * L * ctx * c.dim
        return float(c.num_layers * per)
