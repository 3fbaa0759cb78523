"""LongCat-Video DiT forward on the sm_100a kernels: the object that stands in for ``pipe.dit``.

Call surface of the reference pipeline (longcat_for_worldforge/longcat_video/pipeline_longcat_video.py:867-873):
``dit(hidden_states[B,16,T,H,W], timestep[B,T], encoder_hidden_states[B,1,N,4096], encoder_attention_mask[B,N],
num_cond_latents=1) -> fp32 Tensor[B,16,T,H,W]``, plus ``.dtype`` (:725), ``.config.in_channels`` (:773),
``.cp_split_hw`` (:689).

The arithmetic is ``LongCatVideoTransformer3DModel`` (longcat_video/modules/longcat_video_dit.py:280-370) in the
reference's GPU configuration - bf16 module, fp32 islands (see oracle/longcat_dit.py): bf16 residual stream, per-FRAME
adaLN modulation (timestep [B,T], the condition frame at t = 0), per-head RMSNorm with a bf16 gain, fp32 rotate-half RoPE,
condition / noise split self-attention (attention.py:124-135), cross-attention for the noise tokens only (:262-273),
SwiGLU feed-forward.  It reuses the Wan kernels (tcgen05 GEMM with a bf16-residual epilogue and per-frame gates,
tcgen05 flash attention on column slices of the fused qkv buffer, fused LayerNorm+modulate) plus four LongCat-specific
ones (``wf_rms_norm_head_rope``, ``wf_swiglu_bf16``, ``wf_timestep_embedding_f32``, ``wf_small_gemm_f32``).

LoRA (longcat_video_dit.py:197-249): the reference applies ``W x + m*a*up(down(x))`` at run time; merge it into the weights
before ``from_state_dict`` (``W' = W + m*a*up@down``) - not done here.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Tuple

import torch

from . import lib
from .ctx_cache import ContextCache

BF, F32 = torch.bfloat16, torch.float32


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
    def ffn_dim(self):
        h = int(2 * int(self.hidden_size * self.mlp_ratio) / 3)
        return 256 * ((h + 255) // 256)


def rope_table(grid, head_dim: int = 128) -> torch.Tensor:
    """fp32 [N, head_dim/2, 2] (cos, sin) of rope_3d.py:63-95,109-112: fp32 angles position*freq, then cos / sin in fp32.
    The reference repeats every frequency for the two members of a pair, so one (cos, sin) per pair suffices."""
    T, H, W = grid
    d6 = head_dim // 6
    parts = []
    for n, dim in zip((T, H, W), (head_dim - 4 * d6, 2 * d6, 2 * d6)):
        f = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        parts.append(torch.arange(n, dtype=F32)[:, None] * f[None])
    ft, fh, fw = parts
    a = torch.cat([ft[:, None, None, :].expand(T, H, W, -1), fh[None, :, None, :].expand(T, H, W, -1),
                   fw[None, None, :, :].expand(T, H, W, -1)], dim=-1).reshape(T * H * W, head_dim // 2)
    return torch.stack([a.cos(), a.sin()], dim=-1).contiguous()


def merge_lora(sd: Dict[str, torch.Tensor], lora_sd: Dict[str, torch.Tensor], multiplier: float = 1.0, rank: int = 128,
               alpha: float = 64.0) -> Dict[str, torch.Tensor]:
    """Fold a LongCat LoRA (lora_utils.py:27-75; key scheme ``lora___lorahyphen___<module path with ___lorahyphen___>``)
    into the Linear weights it decorates: W' = W + multiplier * alpha/rank * up @ down, ``up`` block-diagonal over
    ``lora_up.blocks.i`` for the fused qkv / kv projections (LoRAUPParallel, :15-24).  The reference evaluates the LoRA
    branch as two bf16 side GEMMs at run time (longcat_video_dit.py:234-249); folding differs from that by bf16
    rounding of the branch only and removes the side GEMMs from every step of the refine pass."""
    out = dict(sd)
    scale = multiplier * (alpha / rank if alpha else 1.0)
    names = sorted(k[: -len(".lora_down.weight")] for k in lora_sd if k.endswith(".lora_down.weight"))
    for name in names:
        module = name.replace("lora___lorahyphen___", "").replace("___lorahyphen___", ".")
        wkey = module + ".weight"
        if wkey not in out:
            continue
        down = lora_sd[name + ".lora_down.weight"].float()                     # [n*rank, in]
        if name + ".lora_up.weight" in lora_sd:
            delta = lora_sd[name + ".lora_up.weight"].float() @ down
        else:
            blocks = []
            i = 0
            while f"{name}.lora_up.blocks.{i}.weight" in lora_sd:
                blocks.append(lora_sd[f"{name}.lora_up.blocks.{i}.weight"].float())
                i += 1
            r = down.shape[0] // len(blocks)
            delta = torch.cat([u @ down[j * r:(j + 1) * r] for j, u in enumerate(blocks)], dim=0)
        w = out[wkey]
        out[wkey] = (w.float() + scale * delta.to(w.device)).to(w.dtype)
    return out


class WfLongCatTransformer:
    def __init__(self, cfg: LongCatConfig, device):
        assert cfg.hidden_size // cfg.num_heads == 128 and cfg.patch == (1, 2, 2)
        self.cfg, self.device = cfg, torch.device(device)
        self.dtype = BF
        self.config = SimpleNamespace(in_channels=cfg.in_channels, out_channels=cfg.out_channels, patch_size=cfg.patch)
        self.cp_split_hw = [1, 1]
        self.blocks = []
        self._buf, self._rope, self._ctx_cache = {}, {}, ContextCache(4)
        self.calls = 0
        self.bsa_params = None          # the checkpoint's bsa_params (longcat_video_dit.py:32,56)
        self._bsa_on = False
        self.lora_dict, self.active_loras, self._lora_saved = {}, [], {}
        self.cp = None                  # enable_context_parallel

    def to(self, *a, **k):
        return self

    # ---------------------------------------------------------------------------------- LoRA surface
    # longcat_video_dit.py:197-270, used by run_longcat_worldforge_single.py:213-214 (cfg_step_lora of the distilled mode)
    # and :449-451 / :490 (refinement_lora around the 720p pass).  The reference hooks two bf16 side GEMMs into every
    # decorated Linear at run time; here enable_loras FOLDS the active LoRAs into the weights those Linears own
    # (W' = bf16(W + sum multiplier * alpha/rank * up @ down), exactly ``merge_lora``) and disable_all_loras restores the
    # saved originals - the step itself runs the same kernels with or without a LoRA.
    def _lora_targets(self):
        """module path (as the LoRA keys spell it) -> (tensor that holds the weight, first row, rows, fp32 island?)"""
        c = self.cfg
        C, Fd = c.hidden_size, c.ffn_dim
        t = {"final_layer.linear": (self.final_w, 0, self.final_w.shape[0], True),
             "final_layer.adaLN_modulation.1": (self.ada_w, 6 * C * c.depth, 2 * C, False),
             "t_embedder.mlp.0": (self.t0_w, 0, self.t0_w.shape[0], False), "t_embedder.mlp.2": (self.t2_w, 0, self.t2_w.shape[0], False),
             "y_embedder.y_proj.0": (self.y0_w, 0, self.y0_w.shape[0], False), "y_embedder.y_proj.2": (self.y2_w, 0, self.y2_w.shape[0], False)}
        for i, b in enumerate(self.blocks):
            p = f"blocks.{i}."
            t[p + "attn.qkv"] = (b.qkv_w, 0, 3 * C, False)
            t[p + "attn.proj"] = (b.proj_w, 0, C, False)
            t[p + "cross_attn.q_linear"] = (b.cq_w, 0, C, False)
            t[p + "cross_attn.kv_linear"] = (b.ckv_w, 0, 2 * C, False)
            t[p + "cross_attn.proj"] = (b.cproj_w, 0, C, False)
            t[p + "ffn.w1"] = (b.w13, 0, Fd, False)
            t[p + "ffn.w3"] = (b.w13, Fd, Fd, False)
            t[p + "ffn.w2"] = (b.w2, 0, C, False)
            t[p + "adaLN_modulation.1"] = (self.ada_w, 6 * C * i, 6 * C, False)
        return t

    def load_lora(self, lora_path, lora_key, multiplier=1.0, lora_network_dim=128, lora_network_alpha=64):
        """``lora_path``: a .safetensors file (as in the reference) or an already loaded dict of tensors."""
        if isinstance(lora_path, dict):
            sd = lora_path
        else:
            from safetensors.torch import load_file
            sd = load_file(lora_path, device="cpu")
        self.lora_dict[lora_key] = SimpleNamespace(sd=sd, multiplier=multiplier, dim=lora_network_dim, alpha=lora_network_alpha)

    def enable_loras(self, lora_key_list=()):
        self.disable_all_loras()
        targets = self._lora_targets()
        deltas = {}
        for key in lora_key_list:
            if key not in self.lora_dict:
                continue
            net = self.lora_dict[key]
            scale = net.multiplier * ((net.alpha / net.dim) if net.alpha else 1.0)
            names = sorted(k[: -len(".lora_down.weight")] for k in net.sd if k.endswith(".lora_down.weight"))
            for name in names:
                module = name.replace("lora___lorahyphen___", "").replace("___lorahyphen___", ".")
                if module not in targets:
                    raise lib.WfError(f"LoRA '{key}' decorates '{module}', which is not a Linear of this engine")
                down = net.sd[name + ".lora_down.weight"].to(self.device, F32)
                if name + ".lora_up.weight" in net.sd:
                    delta = net.sd[name + ".lora_up.weight"].to(self.device, F32) @ down
                else:                                  # LoRAUPParallel (lora_utils.py:15-24): block-diagonal up over the fused outputs
                    ups = []
                    while f"{name}.lora_up.blocks.{len(ups)}.weight" in net.sd:
                        ups.append(net.sd[f"{name}.lora_up.blocks.{len(ups)}.weight"].to(self.device, F32))
                    r = down.shape[0] // len(ups)
                    delta = torch.cat([u @ down[j * r:(j + 1) * r] for j, u in enumerate(ups)], dim=0)
                deltas[module] = deltas.get(module, 0) + scale * delta
            self.active_loras.append(key)
        for module, delta in deltas.items():
            w, r0, n, island = targets[module]
            view = w[r0:r0 + n]
            self._lora_saved[module] = view.clone()
            view.copy_((view.to(F32) + delta).to(BF).to(view.dtype) if island else (view.to(F32) + delta).to(view.dtype))
        if deltas:
            self._ctx_cache = ContextCache(4)           # cached context K/V came from the weights that just changed

    def disable_all_loras(self):
        targets = self._lora_targets() if self._lora_saved else {}
        for module, saved in self._lora_saved.items():
            w, r0, n, _ = targets[module]
            w[r0:r0 + n].copy_(saved)
        if self._lora_saved:
            self._ctx_cache = ContextCache(4)
        self._lora_saved = {}
        self.active_loras = []

    # ---------------------------------------------------------------------------------- context parallel (2-D split)
    def enable_context_parallel(self, group, split_hw):
        """LongCat's context parallel (longcat_video_dit.py:329-332,361-362; context_parallel_util.py:91-121,180-243;
        ulysses_wrapper.py:87-105): the patch grid of every frame is cut into split_h x split_w blocks, rank r = ih * split_w +
        iw keeps block (ih, iw) of ALL frames; every token-local op runs on the block, self-attention exchanges heads for
        tokens (one all-to-all per q / k / v and one back), and the gathered sequence is rank-major - P copies of a
        (T, H/split_h, W/split_w) volume - which is the geometry the block-sparse attention chunks in (attention.py:60-66), so
        its selections equal the reference's at the same split.  The final layer's output blocks are gathered back."""
        import torch.distributed as dist
        from . import ulysses
        sh, sw = split_hw
        world = dist.get_world_size(group)
        if sh * sw != world:
            raise lib.WfError(f"cp_split_hw {sh} x {sw} does not match the {world} ranks of the group")
        if self.cfg.num_heads % world:
            raise lib.WfError(f"{self.cfg.num_heads} heads do not split over {world} ranks")
        self.cp = SimpleNamespace(group=group, world=world, rank=dist.get_rank(group), split_hw=(sh, sw),
                                  sp=ulysses.GeneralSequenceParallel(group))
        self.cp_split_hw = [sh, sw]
        return self.cp

    # block-sparse self-attention of the 720p refine pass (longcat_video_dit.py:272-278, attention.py:56-67)
    def enable_bsa(self):
        if self.bsa_params is None:
            raise lib.WfError("enable_bsa(): no bsa_params (set .bsa_params to the checkpoint's dict first)")
        if self.bsa_params.get("sparsity", 0.875) is None and self.bsa_params.get("cdf_threshold") is None:
            raise lib.WfError("enable_bsa(): bsa_params needs a sparsity and / or a cdf_threshold (bsa_interface.py:261-270)")
        self._bsa_on = True

    def disable_bsa(self):
        self._bsa_on = False

    def _self_attention(self, q, k, v, out, grid_q, grid_k, sparse: bool, heads: int = None):
        """One _process_attn call (attention.py:49-103): dense, or gating + block-sparse when BSA is enabled.  ``heads``: the
        heads present in q / k / v (all of them, or this rank's share under context parallel)."""
        Hn = self.cfg.num_heads if heads is None else heads
        if not sparse:
            return lib.attention_bf16(q, k, v, out, Hn)
        bp = self.bsa_params
        chunk = tuple(bp.get("chunk_3d_shape_q", (4, 4, 8)))
        if tuple(bp.get("chunk_3d_shape_k", (4, 4, 8))) != chunk:
            raise lib.WfError("BSA with different query / key chunk shapes is not implemented")
        q_cmp = lib.bsa_mean_pool(q, grid_q, chunk, Hn)
        k_cmp = lib.bsa_mean_pool(k, grid_k, chunk, Hn)
        sparsity, cdf = bp.get("sparsity", 0.875), bp.get("cdf_threshold")      # flash_attn_bsa_3d's defaults (bsa_interface.py:619)
        n_sel = int((1 - sparsity) * k_cmp.shape[1]) if sparsity is not None else 0   # :223
        if cdf is not None:                                                    # softmax-mass rule, floored by the top-k count (:234-275)
            idx, lens = lib.bsa_select_cdf(q_cmp, k_cmp, cdf, n_sel)
            return lib.attention_bsa_bf16(q, k, v, out, Hn, idx, lens, grid_q, grid_k, chunk)
        if n_sel < 1:
            out.zero_()                                                        # nothing selected: the kernel's acc = 0, l = 1
            return out
        idx = lib.bsa_select_topk(q_cmp, k_cmp, n_sel)
        return lib.attention_bsa_bf16(q, k, v, out, Hn, idx, None, grid_q, grid_k, chunk)

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], cfg: LongCatConfig, device) -> "WfLongCatTransformer":
        self = cls(cfg, device)
        dev = self.device
        bf = lambda k: sd[k].to(device=dev, dtype=BF).contiguous()
        f32_of_bf = lambda k: sd[k].to(device=dev, dtype=BF).to(F32).contiguous()       # bf16 parameter used in an fp32 island
        C = cfg.hidden_size
        self.patch_w = sd["x_embedder.proj.weight"].flatten(1).to(device=dev, dtype=BF).contiguous()
        self.patch_b = bf("x_embedder.proj.bias")
        self.t0_w, self.t0_b, self.t2_w, self.t2_b = bf("t_embedder.mlp.0.weight"), bf("t_embedder.mlp.0.bias"), bf("t_embedder.mlp.2.weight"), bf("t_embedder.mlp.2.bias")
        self.y0_w, self.y0_b, self.y2_w, self.y2_b = bf("y_embedder.y_proj.0.weight"), bf("y_embedder.y_proj.0.bias"), bf("y_embedder.y_proj.2.weight"), bf("y_embedder.y_proj.2.bias")
        self.final_w, self.final_b = f32_of_bf("final_layer.linear.weight"), f32_of_bf("final_layer.linear.bias")
        ada_w, ada_b = [], []
        for i in range(cfg.depth):
            p = f"blocks.{i}."
            b = SimpleNamespace()
            b.qkv_w, b.qkv_b = bf(p + "attn.qkv.weight"), bf(p + "attn.qkv.bias")
            b.qn, b.kn = bf(p + "attn.q_norm.weight"), bf(p + "attn.k_norm.weight")
            b.proj_w, b.proj_b = bf(p + "attn.proj.weight"), bf(p + "attn.proj.bias")
            b.cq_w, b.cq_b = bf(p + "cross_attn.q_linear.weight"), bf(p + "cross_attn.q_linear.bias")
            b.ckv_w, b.ckv_b = bf(p + "cross_attn.kv_linear.weight"), bf(p + "cross_attn.kv_linear.bias")
            b.cproj_w, b.cproj_b = bf(p + "cross_attn.proj.weight"), bf(p + "cross_attn.proj.bias")
            b.cqn, b.ckn = bf(p + "cross_attn.q_norm.weight"), bf(p + "cross_attn.k_norm.weight")
            b.n_w, b.n_b = f32_of_bf(p + "pre_crs_attn_norm.weight"), f32_of_bf(p + "pre_crs_attn_norm.bias")
            b.w13 = torch.cat([sd[p + "ffn.w1.weight"], sd[p + "ffn.w3.weight"]], dim=0).to(device=dev, dtype=BF).contiguous()
            b.w2 = bf(p + "ffn.w2.weight")
            self.blocks.append(b)
            ada_w.append(sd[p + "adaLN_modulation.1.weight"]); ada_b.append(sd[p + "adaLN_modulation.1.bias"])
        ada_w.append(sd["final_layer.adaLN_modulation.1.weight"]); ada_b.append(sd["final_layer.adaLN_modulation.1.bias"])
        self.ada_w = torch.cat(ada_w, dim=0).to(device=dev, dtype=BF).contiguous()      # [(6*depth + 2)*C, A]
        self.ada_b = torch.cat(ada_b, dim=0).to(device=dev, dtype=BF).contiguous()
        return self

    @classmethod
    def random_init(cls, cfg: LongCatConfig, device, seed: int = 1234) -> "WfLongCatTransformer":
        """Random-init weights generated on the device (LongCat-Video is 13.6 B parameters, 27 GB in bf16): N(0, 0.02^2)
        matrices and biases, norm gains 1 + N(0, 0.05^2) - the same rule as the Wan benchmark model."""
        self = cls(cfg, device)
        g = torch.Generator(device=self.device).manual_seed(seed)
        C, Fd, A = cfg.hidden_size, cfg.ffn_dim, cfg.adaln_tembed_dim
        def m(n, k, dt=BF):
            return (torch.randn(n, k, generator=g, device=self.device, dtype=F32) * 0.02).to(dt)
        def v(n, dt=BF, s=0.02, base=0.0):
            return (base + torch.randn(n, generator=g, device=self.device, dtype=F32) * s).to(dt)
        self.patch_w, self.patch_b = m(C, cfg.in_channels * 4), v(C)
        self.t0_w, self.t0_b, self.t2_w, self.t2_b = m(A, cfg.frequency_embedding_size), v(A), m(A, A), v(A)
        self.y0_w, self.y0_b, self.y2_w, self.y2_b = m(C, cfg.caption_channels), v(C), m(C, C), v(C)
        self.final_w, self.final_b = m(cfg.out_channels * 4, C).to(F32), v(cfg.out_channels * 4).to(F32)
        for _ in range(cfg.depth):
            b = SimpleNamespace()
            b.qkv_w, b.qkv_b = m(3 * C, C), v(3 * C)
            b.qn, b.kn = v(128, BF, 0.05, 1.0), v(128, BF, 0.05, 1.0)
            b.proj_w, b.proj_b = m(C, C), v(C)
            b.cq_w, b.cq_b, b.ckv_w, b.ckv_b = m(C, C), v(C), m(2 * C, C), v(2 * C)
            b.cproj_w, b.cproj_b = m(C, C), v(C)
            b.cqn, b.ckn = v(128, BF, 0.05, 1.0), v(128, BF, 0.05, 1.0)
            b.n_w, b.n_b = v(C, BF, 0.05, 1.0).to(F32), v(C).to(F32)
            b.w13, b.w2 = m(2 * Fd, C), m(C, Fd)
            self.blocks.append(b)
        self.ada_w, self.ada_b = m((6 * cfg.depth + 2) * C, A), v((6 * cfg.depth + 2) * C)
        return self

    def flops_per_forward(self, N: int, n_cond: int, ctx: int) -> float:
        """Algorithmic FLOPs of one forward over N tokens of which n_cond are condition tokens (2 per MAC): qkv / proj / SwiGLU
        on every token, cross-attention on the noise tokens, self-attention cond->cond and noise->all."""
        c = self.cfg
        C, Fd, Nn = c.hidden_size, c.ffn_dim, N - n_cond
        per = 2 * N * (4 * C * C + 3 * C * Fd) + 2 * Nn * 2 * C * C + 2 * ctx * 2 * C * C + 4 * C * (n_cond * n_cond + Nn * N) + 4 * Nn * ctx * C
        return float(c.depth * per)

    def _buffers(self, N, Nn, T):
        key = (N, Nn, T)
        if key not in self._buf:
            c, dev = self.cfg, self.device
            C, Fd = c.hidden_size, c.ffn_dim
            e = lambda *s, dt=BF: torch.empty(*s, dtype=dt, device=dev)
            self._buf[key] = SimpleNamespace(
                cols=e(N, c.in_channels * 4), x=e(N, C), h=e(N, C), qkv=e(N, 3 * C), att=e(N, C), cq=e(Nn, C), ca=e(Nn, C),
                h13=e(N, 2 * Fd), ff=e(N, Fd), temb=e(T, c.frequency_embedding_size, dt=F32), t1=e(T, c.adaln_tembed_dim, dt=F32),
                t=e(T, c.adaln_tembed_dim, dt=F32), mod=e(T, (6 * c.depth + 2) * C, dt=F32))
        return self._buf[key]

    def _context(self, ctx):
        """Caption embedding and every block's cross-attention K|V: functions of the prompt only, computed once."""
        hit = self._ctx_cache.get((ctx,))                     # entries hold their source tensor: see ctx_cache.py
        if hit is not None:
            return hit
        c, dev = self.cfg, self.device
        M, C = ctx.shape[0], c.hidden_size
        e = lambda *s: torch.empty(*s, dtype=BF, device=dev)
        y1 = lib.gemm_bf16(ctx.to(BF).contiguous(), self.y0_w, self.y0_b, e(M, C), lib.EPI_GELU_BF16)
        y = lib.gemm_bf16(y1, self.y2_w, self.y2_b, e(M, C), lib.EPI_BF16)
        kvs = []
        for b in self.blocks:
            kv = lib.gemm_bf16(y, b.ckv_w, b.ckv_b, e(M, 2 * C), lib.EPI_BF16)
            # kv_linear output is [M, 2, H, 128] (attention.py:215): k = first C columns, v = last C
            lib.rms_norm_head_rope_(kv[:, :C], b.ckn, 1e-6, None)
            kvs.append(kv)
        self._ctx_cache.put((ctx,), kvs)
        return kvs

    @torch.no_grad()
    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask=None, num_cond_latents: int = 0,
                 return_kv: bool = False, kv_cache_dict=None, skip_crs_attn: bool = False, offload_kv_cache: bool = False, **kw):
        """``return_kv`` / ``kv_cache_dict`` / ``skip_crs_attn`` / ``offload_kv_cache``: the video-continuation surface of the
        reference forward (longcat_video_dit.py:280-370; pipeline_longcat_video.py:336-350, 1196-1209).  With return_kv the
        call returns (out, {layer: (k, v)}): per layer the self-attention keys / values of the given (clean condition) frames,
        bf16 [B, tokens, C] - an opaque cache for THIS engine: keys are held before the q/k norm and RoPE (the reference caches
        them after the norm; the fused norm+RoPE kernel re-applies both over [cache | own] when the cache is used, which is
        the same arithmetic).  With a kv_cache_dict, hidden_states holds the noise frames only and num_cond_latents is the
        number of cached frames."""
        if not hidden_states.is_cuda:
            raise lib.WfError("WfLongCatTransformer runs on CUDA tensors only (no CPU fallback)")
        B = hidden_states.shape[0]
        if timestep.dim() == 1:
            timestep = timestep.unsqueeze(1).expand(-1, hidden_states.shape[2])
        if return_kv or kv_cache_dict:
            outs, caches = [], []
            for s in range(B):
                ctx = encoder_hidden_states[s, 0]
                if encoder_attention_mask is not None and not skip_crs_attn:
                    m = encoder_attention_mask[s].reshape(-1) != 0
                    ctx = ctx[m] if not bool(m[:int(m.sum())].all()) else ctx[:int(m.sum())]
                cache_s = None
                if kv_cache_dict:
                    cache_s = {i: (kv[0][min(s, kv[0].shape[0] - 1)].to(self.device), kv[1][min(s, kv[1].shape[0] - 1)].to(self.device))
                               for i, kv in kv_cache_dict.items()}
                r = self._forward_one(hidden_states[s], timestep[s], ctx, 0 if kv_cache_dict else num_cond_latents,
                                      return_kv=return_kv, kv_cache=cache_s, cached_frames=num_cond_latents if kv_cache_dict else 0,
                                      skip_crs=skip_crs_attn)
                if return_kv:
                    outs.append(r[0]); caches.append(r[1])
                else:
                    outs.append(r)
            out = torch.stack(outs)
            if not return_kv:
                return out
            ret = {}
            for i in caches[0]:
                k = torch.stack([c[i][0] for c in caches]); v = torch.stack([c[i][1] for c in caches])
                ret[i] = (k.cpu(), v.cpu()) if offload_kv_cache else (k, v)
            return out, ret
        outs = []
        for s in range(B):
            ctx = encoder_hidden_states[s, 0]
            if encoder_attention_mask is not None:
                m = encoder_attention_mask[s].reshape(-1) != 0
                n_valid = int(m.sum())
                if not bool(m[:n_valid].all()):
                    ctx = ctx[m]                       # general mask: gather the valid tokens (longcat_video_dit.py:323)
                else:
                    ctx = ctx[:n_valid]
            outs.append(self._forward_one(hidden_states[s], timestep[s], ctx, num_cond_latents))
        return torch.stack(outs)

    def _forward_one(self, x, timestep, ctx, num_cond, return_kv=False, kv_cache=None, cached_frames=0, skip_crs=False):
        c = self.cfg
        self.calls += 1
        C, Hn, Fd = c.hidden_size, c.num_heads, c.ffn_dim
        xb = x.to(BF).contiguous()
        _, T, H, W = xb.shape
        full_grid = (T, H // 2, W // 2)
        cp = self.cp
        if cp is not None:
            if kv_cache is not None or return_kv:
                raise lib.WfError("context parallel together with the KV cache is not supported")
            from . import ulysses
            sh, sw = cp.split_hw
            if full_grid[1] % sh or full_grid[2] % sw:
                raise lib.WfError(f"patch grid {full_grid[1]} x {full_grid[2]} is not a multiple of cp_split_hw {sh} x {sw}")
            grid = (T, full_grid[1] // sh, full_grid[2] // sw)          # this rank's block of every frame
        else:
            grid = full_grid
        per = grid[1] * grid[2]
        N = T * per
        nc = num_cond * per
        Nn = N - nc
        Bf = self._buffers(N, Nn, T)
        rgrid = (T + cached_frames, full_grid[1], full_grid[2])  # RoPE runs over [cached | own] frames (attention.py:171)
        rkey = rgrid if cp is None else (rgrid, cp.split_hw, cp.rank)
        if rkey not in self._rope:
            tab = rope_table(rgrid)
            if cp is not None:                                  # global positions of this rank's block (rope_3d.py:79-86)
                tab = ulysses.split_2d(tab.view(rgrid[0], rgrid[1], rgrid[2], 64, 2), (1, 2), cp.split_hw, cp.rank).reshape(-1, 64, 2).contiguous()
            self._rope[rkey] = tab.to(self.device)
        Nc = cached_frames * per                                # cached condition tokens in front of this call's tokens
        rope_all = self._rope[rkey]
        rope = rope_all[Nc:]
        kv_ret = {}
        if kv_cache is not None:
            kfull = torch.empty(Nc + N, C, dtype=BF, device=self.device)
            vfull = torch.empty(Nc + N, C, dtype=BF, device=self.device)

        if cp is None:
            lib.patchify(xb, Bf.cols)
        else:                                                   # patch columns of the whole clip, then this rank's block
            cols_full = torch.empty(full_grid[0] * full_grid[1] * full_grid[2], Bf.cols.shape[1], dtype=BF, device=self.device)
            lib.patchify(xb, cols_full)
            Bf.cols.copy_(ulysses.split_2d(cols_full.view(*full_grid, -1), (1, 2), cp.split_hw, cp.rank).reshape(N, -1))
        lib.gemm_bf16(Bf.cols, self.patch_w, self.patch_b, Bf.x, lib.EPI_BF16)
        # per-frame timestep embedding and ALL adaLN tables in fp32: timestep.to(bf16).float() as the reference does (:307,:313)
        ts = timestep.to(BF).to(F32).contiguous()
        lib.timestep_embedding_f32(ts, Bf.temb)
        lib.small_gemm_f32(Bf.temb, self.t0_w, self.t0_b, Bf.t1)
        lib.small_gemm_f32(Bf.t1, self.t2_w, self.t2_b, Bf.t, silu_in=True)
        lib.small_gemm_f32(Bf.t, self.ada_w, self.ada_b, Bf.mod, silu_in=True)
        mod = Bf.mod.view(T, 6 * c.depth + 2, C).permute(1, 0, 2).contiguous()      # [table, T, C]: each table contiguous
        kvs = None if skip_crs else self._context(ctx)

        for i, b in enumerate(self.blocks):
            # [T, C] tables of this block: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp (chunk order, :83-85)
            tab = [mod[6 * i + j] for j in range(6)]
            lib.layer_norm(Bf.x, Bf.h, 1e-6, scale=tab[1], shift=tab[0], rows_per_group=per)
            lib.gemm_bf16(Bf.h, b.qkv_w, b.qkv_b, Bf.qkv, lib.EPI_BF16)
            if return_kv:                                   # before the in-place norm + RoPE (see __call__)
                kv_ret[i] = (Bf.qkv[:, C:2 * C].clone(), Bf.qkv[:, 2 * C:].clone())
            lib.rms_norm_head_rope_(Bf.qkv[:, :C], b.qn, 1e-6, rope)
            if kv_cache is not None:                        # forward_with_kv_cache (attention.py:149-181)
                kfull[:Nc].copy_(kv_cache[i][0]); kfull[Nc:].copy_(Bf.qkv[:, C:2 * C])
                vfull[:Nc].copy_(kv_cache[i][1]); vfull[Nc:].copy_(Bf.qkv[:, 2 * C:])
                lib.rms_norm_head_rope_(kfull, b.kn, 1e-6, rope_all)
            else:
                lib.rms_norm_head_rope_(Bf.qkv[:, C:2 * C], b.kn, 1e-6, rope)
            q, k, v = Bf.qkv[:, :C], Bf.qkv[:, C:2 * C], Bf.qkv[:, 2 * C:]
            # in the reference BSA is skipped for single-frame inputs as a whole (shape[0] > 1, attention.py:56)
            sparse = self._bsa_on and T > 1
            gq = lambda t: (t, grid[1], grid[2])
            if cp is not None:
                # Ulysses (ulysses_wrapper.py:87-105): heads <-> tokens; the gathered sequence is P copies of the local volume
                P = cp.world
                gP = lambda t: (P * t, grid[1], grid[2])
                def exchange(qv, kv_, vv, outv, tq, tk):
                    fn = lambda qf, kf, vf, of, hl: self._self_attention(qf, kf, vf, of, gP(tq), gP(tk), sparse, heads=hl)
                    outv.copy_(cp.sp.attention_qkv(qv.contiguous(), kv_.contiguous(), vv.contiguous(), Hn, fn))
                if nc > 0:
                    exchange(q[:nc], k[:nc], v[:nc], Bf.att[:nc], num_cond, num_cond)
                    exchange(q[nc:], k, v, Bf.att[nc:], T - num_cond, T)
                else:
                    exchange(q, k, v, Bf.att, T, T)
            elif kv_cache is not None:
                self._self_attention(q, kfull, vfull, Bf.att, gq(T), gq(T + cached_frames), sparse)
            elif nc > 0:      # condition tokens see only condition tokens; noise tokens see everything (attention.py:124-135)
                self._self_attention(q[:nc], k[:nc], v[:nc], Bf.att[:nc], gq(num_cond), gq(num_cond), sparse)
                self._self_attention(q[nc:], k, v, Bf.att[nc:], gq(T - num_cond), gq(T), sparse)
            else:
                self._self_attention(q, k, v, Bf.att, gq(T), gq(T), sparse)
            lib.gemm_bf16(Bf.att, b.proj_w, b.proj_b, Bf.x, lib.EPI_RESID_BF16, gate=tab[2], gate_rows=per)
            # cross attention, noise tokens only
            if not skip_crs:
                lib.layer_norm(Bf.x, Bf.h, 1e-6, weight=b.n_w, bias=b.n_b)
                lib.gemm_bf16(Bf.h[nc:], b.cq_w, b.cq_b, Bf.cq, lib.EPI_BF16)
                lib.rms_norm_head_rope_(Bf.cq, b.cqn, 1e-6, None)
                kv = kvs[i]
                lib.attention_bf16(Bf.cq, kv[:, :C], kv[:, C:], Bf.ca, Hn)
                lib.gemm_bf16(Bf.ca, b.cproj_w, b.cproj_b, Bf.x[nc:], lib.EPI_RESID_BF16)
            # SwiGLU feed-forward
            lib.layer_norm(Bf.x, Bf.h, 1e-6, scale=tab[4], shift=tab[3], rows_per_group=per)
            lib.gemm_bf16(Bf.h, b.w13, None, Bf.h13, lib.EPI_BF16)
            lib.swiglu_bf16(Bf.h13, Bf.ff)
            lib.gemm_bf16(Bf.ff, b.w2, None, Bf.x, lib.EPI_RESID_BF16, gate=tab[5], gate_rows=per)

        shift, scale = mod[6 * c.depth], mod[6 * c.depth + 1]
        out = torch.empty(c.out_channels, T, 2 * grid[1], 2 * grid[2], dtype=F32, device=self.device)
        lib.dit_head(Bf.x, scale, shift, self.final_w, self.final_b, out, grid, 1e-6, rows_per_group=per, round_bf16=True)
        if cp is not None:                                      # gather_cp_2d (longcat_video_dit.py:361-362)
            import torch.distributed as dist
            parts = [torch.empty_like(out) for _ in range(cp.world)]
            dist.all_gather(parts, out, group=cp.group)
            out = ulysses.gather_2d(parts, (2, 3), cp.split_hw).contiguous()
        return (out, kv_ret) if return_kv else out
