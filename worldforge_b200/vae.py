"""Wan2.1 causal 3D-VAE on the sm_100a kernels: the object that stands in for ``pipe.vae``.

Surface used by the reference: ``encode(x[1,3,F,H,W]).latent_dist.mode()``
(pipeline_wan_i2v_clean.py:348, scheduling_unipc_multistep_clean.py:1384), ``decode(z, return_dict=False)[0]``
(:743, :1285), ``config.{z_dim, latents_mean, latents_std}``, ``temperal_downsample``, ``dtype``.

The network is the reference's ``WanVAE_`` (wan/modules/vae.py) = diffusers' ``AutoencoderKLWan``.  The reference
walks the clip chunk by chunk with a per-convolution feature cache (vae.py:516-568); the engine evaluates every
layer over the whole clip at once (the same function - see oracle/wan_vae.py for the three identities used and
their numerical check against the chunked reference), which turns 21 x 33 tiny cuDNN launches per decode into a
few dozen full-grid launches:

* activations are channels-last fp32 ``[T][H][W][C]`` so that a convolution tap is a shifted 4-D TMA box;
* every convolution is ``wf_conv_tf32`` (tcgen05, tf32 operands = cuDNN's default precision for the reference's
  fp32 VAE, fp32 accumulation).  The tensor core TRUNCATES fp32 operands to tf32 where cuDNN rounds to nearest
  (measured: 7.7e-4 against 2.9e-4 relative error per convolution), so every convolution operand is rounded to tf32
  where it is produced - weights at load, activations by the RMS-norm kernel, the layout converter, a producing
  convolution's epilogue when all its consumers are convolutions (``_round_out``), or ``wf_round_tf32`` for the raw
  residual stream read by the three 1x1 shortcut convolutions - and the truncation is then exact; bias, the residual add of ResidualBlock / AttentionBlock, upsample3d's
  channel->frame de-interleave and the final clamp + planar store are fused into its epilogue;
* nearest-exact 2x upsampling is never materialised: upsample + 3x3 conv = four 2x2-tap convolutions on the
  low-resolution tensor (one per output parity) with pre-summed weights;
* pad(0,1,0,1) + stride-2 3x3 conv = a 2x2-tap convolution on the space-to-depth tensor;
* RMS-norm + SiLU rides in the epilogue of the convolution that PRODUCES its input wherever one accumulator tile holds
  all channels of a pixel (<= 192 channels: every full- and half-resolution layer): the norm between a residual block's two
  convolutions replaces the first one's raw result, and the norm that opens the next block (or the head) is stored next to
  the second one's result; only the 384-channel layers and the layers behind a resampling step use ``wf_rms_norm_cl``.

Multi-GPU (``enable_row_sharding``): the causal feature caches forbid a split along frames, but every layer outside
the mid block is local in space, so the P ranks of one box split the image ROWS.  Each rank evaluates the sharded
part of the network on its rows plus the halo its outputs depend on (found by walking the layer plan backwards:
+1 row per 3x3 conv on each side, halved/doubled by the resampling layers) - the halo is recomputed, not exchanged,
so the convolution kernels are untouched and per-pixel arithmetic is identical to the single-GPU evaluation (the
stitched result is bit-identical).  The mid block (full-frame attention at 1/8 resolution, ~5 % of the work) is
replicated.  One all-gather per encode (1/8-resolution features) and per decode (the decoded rows).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Tuple

import torch

from . import lib

F32 = torch.float32

LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]   # reference vae.py:629-632
LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
               3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]        # reference vae.py:633-636

TAPS_333 = [(dt - 2, dy - 1, dx - 1) for dt in range(3) for dy in range(3) for dx in range(3)]
TAPS_1 = [(0, 0, 0)]
TAPS_T_CAUSAL = [(-2, 0, 0), (-1, 0, 0), (0, 0, 0)]
TAPS_T_STRIDE2 = [(0, 0, 0), (1, 0, 0), (2, 0, 0)]
TAPS_S2D = [(0, a, b) for a in range(2) for b in range(2)]


def _round_tf32_host(w: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest tf32 (ties away from zero, like cvt.rna.tf32.f32), any device."""
    return ((w.contiguous().view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)


def _pad_cin(w2: torch.Tensor) -> torch.Tensor:
    """[rows, Cin] -> Cin padded to a multiple of 4 (16-byte pixel rows for TMA)."""
    cin = w2.shape[1]
    pad = (-cin) % 4
    if pad:
        w2 = torch.cat([w2, w2.new_zeros(w2.shape[0], pad)], dim=1)
    return w2.contiguous()


def _w_conv3d(w: torch.Tensor) -> torch.Tensor:
    """[Co, Ci, kt, kh, kw] -> [(kt*kh*kw)*Co, Ci] tap-major."""
    co, ci = w.shape[:2]
    return _pad_cin(w.permute(2, 3, 4, 0, 1).reshape(-1, ci))


def _w_upsample_parity(w: torch.Tensor) -> List[torch.Tensor]:
    """Conv2d weight [Co, Ci, 3, 3] applied after nearest-exact 2x upsampling == for output parity (p, q) a 2x2-tap
    conv on the low-res input: rows p=0 use (y-1: W[0], y: W[1]+W[2]); p=1 use (y: W[0]+W[1], y+1: W[2])."""
    def comb(t, par):      # t: [..., 3] along one kernel axis -> [..., 2]
        a, b, c = t.unbind(-1)
        return torch.stack([a, b + c], -1) if par == 0 else torch.stack([a + b, c], -1)
    out = []
    for p in (0, 1):
        for q in (0, 1):
            wy = comb(w.transpose(2, 3), p).transpose(2, 3)      # combine over kh -> [Co,Ci,2,3]
            wyx = comb(wy, q)                                     # combine over kw -> [Co,Ci,2,2]
            co, ci = w.shape[:2]
            out.append(_pad_cin(wyx.permute(2, 3, 0, 1).reshape(-1, ci)))
    return out


def _taps_upsample(p: int, q: int):
    ys = (-1, 0) if p == 0 else (0, 1)
    xs = (-1, 0) if q == 0 else (0, 1)
    return [(0, y, x) for y in ys for x in xs]


def _w_s2d(w: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [Co, Ci, 3, 3], stride 2 after pad(0,1,0,1) -> 2x2-tap weights over the space-to-depth input
    whose channel index is (p*2+q)*Ci + c:  W'[a,b][co][(p,q,c)] = W[co,c,2a+p,2b+q] (zero where 2a+p or 2b+q is 3)."""
    co, ci = w.shape[:2]
    out = w.new_zeros(2, 2, co, 4 * ci)
    for a in range(2):
        for b in range(2):
            for p in range(2):
                for q in range(2):
                    if 2 * a + p <= 2 and 2 * b + q <= 2:
                        out[a, b, :, (p * 2 + q) * ci:(p * 2 + q + 1) * ci] = w[:, :, 2 * a + p, 2 * b + q]
    return out.reshape(4 * co, 4 * ci).contiguous()


def assemble_rows(parts: List[torch.Tensor], dim: int) -> torch.Tensor:
    """Stitch the per-rank row slabs (in rank order) back into one tensor."""
    return torch.cat(parts, dim=dim)


class _Dist:
    def __init__(self, mean):
        self._mean = mean

    def mode(self):
        return self._mean

    @property
    def mean(self):
        return self._mean


def to_vendored_vae_names(sd: Dict[str, torch.Tensor], num_res_blocks: int = 2) -> Dict[str, torch.Tensor]:
    """Accept the state dict of diffusers' ``AutoencoderKLWan`` - the class the entry script loads
    (wan_for_worldforge/infer_worldforge.py:185-189; vendored twin longcat_video/modules/autoencoder_kl_wan.py:505-604
    encoder, :783-870 decoder, :1029-1052 quant convs) - and return it under the key names of the reference's
    ``WanVAE_`` (wan/modules/vae.py), which is what ``WfWanVAE`` consumes.  A dict that already uses them passes through.

      encoder.conv_in / conv_out / norm_out            -> encoder.conv1 / head.2 / head.0
      encoder.down_blocks.N  (resnet | resample)        -> encoder.downsamples.N
      *.mid_block.resnets.{0,1} / attentions.0          -> *.middle.{0,2} / middle.1
      decoder.up_blocks.I.resnets.J / upsamplers.0      -> decoder.upsamples.{I*(R+2)+J} / {I*(R+2)+R+1}   (R = num_res_blocks)
      resnet: norm1 / conv1 / norm2 / conv2 / conv_shortcut -> residual.0 / .2 / .3 / .6 / shortcut
      quant_conv / post_quant_conv                      -> conv1 / conv2
    """
    if not any(k.startswith(("quant_conv.", "post_quant_conv.", "encoder.conv_in.", "decoder.conv_in.")) for k in sd):
        return sd
    res = {"norm1": "residual.0", "conv1": "residual.2", "norm2": "residual.3", "conv2": "residual.6", "conv_shortcut": "shortcut"}
    R = num_res_blocks

    def resnet(rest: str) -> str:
        head, _, tail = rest.partition(".")
        return res[head] + "." + tail if head in res else rest

    out = {}
    for k, v in sd.items():
        p = k.split(".")
        if p[0] == "quant_conv":
            out["conv1." + ".".join(p[1:])] = v
        elif p[0] == "post_quant_conv":
            out["conv2." + ".".join(p[1:])] = v
        elif p[0] in ("encoder", "decoder"):
            side, rest = p[0], p[1:]
            if rest[0] == "conv_in":
                nk = "conv1." + ".".join(rest[1:])
            elif rest[0] == "norm_out":
                nk = "head.0." + ".".join(rest[1:])
            elif rest[0] == "conv_out":
                nk = "head.2." + ".".join(rest[1:])
            elif rest[0] == "mid_block":
                if rest[1] == "resnets":
                    nk = f"middle.{2 * int(rest[2])}." + resnet(".".join(rest[3:]))
                else:                                   # attentions.0
                    nk = "middle.1." + ".".join(rest[3:])
            elif rest[0] == "down_blocks":
                nk = f"downsamples.{rest[1]}." + resnet(".".join(rest[2:]))
            elif rest[0] == "up_blocks":
                i = int(rest[1])
                if rest[2] == "resnets":
                    nk = f"upsamples.{i * (R + 2) + int(rest[3])}." + resnet(".".join(rest[4:]))
                else:                                   # upsamplers.0
                    nk = f"upsamples.{i * (R + 2) + R + 1}." + ".".join(rest[4:])
            else:
                nk = ".".join(rest)
            out[side + "." + nk] = v
        else:
            out[k] = v
    return out


class WfWanVAE:
    """Weights prepared once on the device; ``encode`` / ``decode`` are sequences of C-ABI launches."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, dim: int = 96, z_dim: int = 16,
                 dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2, temporal_downsample=(False, True, True)):
        self.device = torch.device(device)
        self.dtype = F32
        self.z_dim = z_dim
        self.config = SimpleNamespace(z_dim=z_dim, latents_mean=list(LATENTS_MEAN[:z_dim]), latents_std=list(LATENTS_STD[:z_dim]))
        self.temperal_downsample = list(temporal_downsample)
        self.dim, self.dim_mult, self.num_res_blocks = dim, tuple(dim_mult), num_res_blocks
        self.enc_plan, self.dec_plan = self._plans()
        self.spatial_scale = 2 ** (len(self.dim_mult) - 1)
        self.shard = None                  # set by enable_row_sharding
        sd = {k: v.to(device=self.device, dtype=F32) for k, v in to_vendored_vae_names(state_dict, num_res_blocks).items()}
        self.w: Dict[str, torch.Tensor] = {}
        self._prepare(sd)

    def to(self, *a, **k):
        return self

    # ------------------------------------------------------------------------------ architecture
    def _plans(self):
        dims = [self.dim * u for u in (1,) + self.dim_mult]
        enc = [("conv", "encoder.conv1", 3, dims[0])]
        idx = 0
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(self.num_res_blocks):
                enc.append(("res", f"encoder.downsamples.{idx}", cin, cout)); idx += 1
                cin = cout
            if i != len(self.dim_mult) - 1:
                enc.append(("down3d" if self.temperal_downsample[i] else "down2d", f"encoder.downsamples.{idx}", cout, cout)); idx += 1
        c = dims[-1]
        enc += [("res", "encoder.middle.0", c, c), ("attn", "encoder.middle.1", c, c), ("res", "encoder.middle.2", c, c),
                ("head", "encoder.head", c, self.z_dim * 2)]
        ddims = [self.dim * u for u in (self.dim_mult[-1],) + self.dim_mult[::-1]]
        up = tuple(self.temperal_downsample[::-1])
        c = ddims[0]
        dec = [("conv", "decoder.conv1", self.z_dim, c), ("res", "decoder.middle.0", c, c), ("attn", "decoder.middle.1", c, c),
               ("res", "decoder.middle.2", c, c)]
        idx = 0
        for i, (cin, cout) in enumerate(zip(ddims[:-1], ddims[1:])):
            if i in (1, 2, 3):
                cin //= 2
            for _ in range(self.num_res_blocks + 1):
                dec.append(("res", f"decoder.upsamples.{idx}", cin, cout)); idx += 1
                cin = cout
            if i != len(self.dim_mult) - 1:
                dec.append(("up3d" if up[i] else "up2d", f"decoder.upsamples.{idx}", cout, cout // 2)); idx += 1
        dec.append(("head", "decoder.head", ddims[-1], 3))
        return enc, dec

    def _prepare(self, sd):
        W = self.w
        def conv3(name):
            W[name + ".w"] = _w_conv3d(sd[name + ".weight"]); W[name + ".b"] = sd[name + ".bias"].contiguous()
        def gamma(name):
            W[name] = sd[name].reshape(-1).contiguous()
        for plan in (self.enc_plan, self.dec_plan):
            for kind, name, cin, cout in plan:
                if kind == "conv":
                    conv3(name)
                elif kind == "res":
                    gamma(name + ".residual.0.gamma"); conv3(name + ".residual.2")
                    gamma(name + ".residual.3.gamma"); conv3(name + ".residual.6")
                    if cin != cout:
                        conv3(name + ".shortcut")
                elif kind == "attn":
                    gamma(name + ".norm.gamma")
                    wq = sd[name + ".to_qkv.weight"].reshape(3, cin, cin)
                    bq = sd[name + ".to_qkv.bias"].reshape(3, cin)
                    for j, t in enumerate("qkv"):
                        W[f"{name}.{t}.w"] = wq[j].contiguous(); W[f"{name}.{t}.b"] = bq[j].contiguous()
                    W[name + ".proj.w"] = sd[name + ".proj.weight"].reshape(cin, cin).contiguous()
                    W[name + ".proj.b"] = sd[name + ".proj.bias"].contiguous()
                elif kind in ("down2d", "down3d"):
                    W[name + ".s2d.w"] = _w_s2d(sd[name + ".resample.1.weight"]); W[name + ".s2d.b"] = sd[name + ".resample.1.bias"].contiguous()
                    if kind == "down3d":
                        conv3(name + ".time_conv")
                elif kind in ("up2d", "up3d"):
                    for j, wp in enumerate(_w_upsample_parity(sd[name + ".resample.1.weight"])):
                        W[f"{name}.par{j}.w"] = wp
                    W[name + ".par.b"] = sd[name + ".resample.1.bias"].contiguous()
                    if kind == "up3d":
                        conv3(name + ".time_conv")
                elif kind == "head":
                    gamma(name + ".0.gamma"); conv3(name + ".2")
        conv3("conv1"); conv3("conv2")
        for k in list(W):                      # convolution weights are tensor-core operands: rounded to tf32 once, here
            if k.endswith(".w"):
                W[k] = _round_tf32_host(W[k])
        # a layer's output may be STORED rounded to tf32 when every consumer is a convolution: the residual block in
        # front of a resampling layer (Resample reads it through convolutions only, vae.py:98-160)
        self._round_out = set()
        for plan in (self.enc_plan, self.dec_plan):
            for (kind, name, _, _), nxt in zip(plan[:-1], plan[1:]):
                if kind == "res" and nxt[0] in ("up2d", "up3d", "down2d", "down3d"):
                    self._round_out.add(name)

    # ------------------------------------------------------------------------------ layers
    @staticmethod
    def _tile_w(w: int) -> int:
        return 16 if w % 16 == 0 else 8

    def _conv(self, x, wname, taps, cout, *, resid=None, t_out=None, t_stride=1, t_off=0, out=None, t_mul=1, c_split=None,
              round_out=False, norm_gamma=None, norm_out=None, norm_silu=True):
        T, H, Wd, _ = x.shape
        t_out = T if t_out is None else t_out
        if out is None:
            out = torch.empty(t_out * t_mul, H, Wd, cout if c_split is None else c_split, dtype=F32, device=x.device)
        lib.conv_tf32(x, self.w[wname + ".w"], self.w[wname + ".b"], taps, out, T=t_out, H=H, W=Wd, Cout=cout,
                      t_stride=t_stride, t_off=t_off, t_mul=t_mul, c_split=c_split, resid=resid, tile_w=self._tile_w(Wd),
                      round_out=round_out, norm_gamma=norm_gamma, norm_out=norm_out, norm_silu=norm_silu)
        return out

    FUSE_NORM_MAX_C = 192      # wf_conv_tf32 fuses the following RMS-norm when one accumulator tile holds a pixel's channels

    def _res(self, x, name, cin, cout, pre=None, next_gamma=None):
        """ResidualBlock (vae.py:186-220).  ``pre``: silu(norm0(x)) if the layer in front already produced it in its
        epilogue.  ``next_gamma``: gamma of the NEXT layer's first norm - then this block's second convolution stores that
        activation too and (out, activation) is returned.  The norm between the two convolutions is always fused into the
        first one's epilogue (its raw result has no other reader) when the width allows."""
        # the 1x1 shortcut reads the raw residual stream: a tf32-rounded copy is its operand (x itself stays exact)
        h = self._conv(lib.round_tf32(x), name + ".shortcut", TAPS_1, cout) if cin != cout else x
        y = pre if pre is not None else lib.rms_norm_cl(x, self.w[name + ".residual.0.gamma"])
        if cout <= self.FUSE_NORM_MAX_C:
            y = self._conv(y, name + ".residual.2", TAPS_333, cout, norm_gamma=self.w[name + ".residual.3.gamma"])
        else:
            y = self._conv(y, name + ".residual.2", TAPS_333, cout)
            lib.rms_norm_cl(y, self.w[name + ".residual.3.gamma"], out=y)
        if next_gamma is not None:
            act = torch.empty(*y.shape[:3], cout, dtype=F32, device=x.device)
            out = self._conv(y, name + ".residual.6", TAPS_333, cout, resid=h, norm_gamma=next_gamma, norm_out=act)
            return out, act
        return self._conv(y, name + ".residual.6", TAPS_333, cout, resid=h, round_out=name in self._round_out), None

    def _attn(self, x, name, c):
        T, H, Wd, _ = x.shape
        n = lib.rms_norm_cl(x, self.w[name + ".norm.gamma"], silu=False)
        q, k, v = (self._conv(n, f"{name}.{t}", TAPS_1, c) for t in "qkv")
        del n
        hw = H * Wd
        o = torch.empty(T, H, Wd, c, dtype=F32, device=x.device)
        s = torch.empty(hw, hw, dtype=F32, device=x.device)
        vt = torch.empty(c, hw, dtype=F32, device=x.device)
        # The reference computes these two matmuls in fp32 (F.scaled_dot_product_attention on fp32 tensors, vae.py:252-256;
        # torch keeps TF32 off for matmuls).  Three tf32 products of the (hi, lo) splits accumulate to fp32 accuracy:
        # A.B = A_hi.B_hi + A_hi.B_lo + A_lo.B_hi + O(2^-22).  60 GFLOP x 3 per frame: ~1 % of a decode.
        e2 = lambda *sh: (torch.empty(*sh, dtype=F32, device=x.device), torch.empty(*sh, dtype=F32, device=x.device))
        q_hl, k_hl, s_hl, v_hl = e2(hw, c), e2(hw, c), e2(hw, hw), e2(c, hw)
        def mm3(a_hl, b_hl, out, n, round_out=False):   # out[hw, n] = a[hw, K] . b[n, K]^T with both operands split
            a_hi, a_lo = (t.view(1, 1, hw, -1) for t in a_hl)
            ov = out.view(1, 1, hw, n)
            lib.conv_tf32(a_lo, b_hl[0], None, TAPS_1, ov, T=1, H=1, W=hw, Cout=n, tile_w=128)          # small terms first
            lib.conv_tf32(a_hi, b_hl[1], None, TAPS_1, ov, T=1, H=1, W=hw, Cout=n, tile_w=128, resid=ov)
            lib.conv_tf32(a_hi, b_hl[0], None, TAPS_1, ov, T=1, H=1, W=hw, Cout=n, tile_w=128, resid=ov, round_out=round_out)
        for t in range(T):
            lib.split_tf32(q[t].view(hw, c), *q_hl)
            lib.split_tf32(k[t].view(hw, c), *k_hl)
            mm3(q_hl, k_hl, s, hw)
            lib.softmax_rows_(s, c ** -0.5)
            lib.transpose_f32(v[t].view(hw, c), vt)
            lib.split_tf32(s, *s_hl)
            lib.split_tf32(vt, *v_hl)
            mm3(s_hl, v_hl, o[t].view(hw, c), c, round_out=True)      # o feeds the proj convolution only
        del q_hl, k_hl, s_hl, v_hl
        del q, k, v, s, vt
        return self._conv(o, name + ".proj", TAPS_1, c, resid=x)

    def _down(self, x, name, c, temporal):
        T, H, Wd, _ = x.shape
        s = lib.space_to_depth(x)                  # x was stored tf32-rounded by the residual block in front (_round_out)
        temporal = temporal and T > 1
        y = torch.empty(T, H // 2, Wd // 2, c, dtype=F32, device=x.device)
        # down3d: the strided 2-D conv's result feeds the temporal conv (rounded to tf32) - except frame 0, which is also
        # passed through as the first output frame and must stay exact: it is computed a second time, unrounded
        lib.conv_tf32(s, self.w[name + ".s2d.w"], self.w[name + ".s2d.b"], TAPS_S2D, y, T=T, H=H // 2, W=Wd // 2, Cout=c,
                      tile_w=self._tile_w(Wd // 2), round_out=temporal)
        if temporal:
            t2 = (T - 1) // 2
            out = torch.empty(1 + t2, H // 2, Wd // 2, c, dtype=F32, device=x.device)
            lib.conv_tf32(s[:1], self.w[name + ".s2d.w"], self.w[name + ".s2d.b"], TAPS_S2D, out[:1], T=1, H=H // 2, W=Wd // 2,
                          Cout=c, tile_w=self._tile_w(Wd // 2))
            del s
            self._conv(y, name + ".time_conv", TAPS_T_STRIDE2, c, t_out=t2, t_stride=2, out=out[1:])
            return out
        return y

    def _up(self, x, name, c, temporal):
        T, H, Wd, _ = x.shape
        if temporal and T > 1:
            xt = torch.empty(1 + 2 * (T - 1), H, Wd, c, dtype=F32, device=x.device)
            xt[0].copy_(x[0])
            # frames 1.. through the temporal conv with an all-zero history; 2C output channels -> two frames of C
            self._conv(x[1:], name + ".time_conv", TAPS_T_CAUSAL, 2 * c, out=xt[1:], t_mul=2, c_split=c, round_out=True)
            x = xt
            T = x.shape[0]
        out = torch.empty(T, 2 * H, 2 * Wd, c // 2, dtype=F32, device=x.device)
        for p in (0, 1):
            for q in (0, 1):
                lib.conv_tf32(x, self.w[f"{name}.par{p * 2 + q}.w"], self.w[name + ".par.b"], _taps_upsample(p, q), out,
                              T=T, H=H, W=Wd, Cout=c // 2, sy=2, sx=2, oy=p, ox=q, tile_w=self._tile_w(Wd))
        return out

    def _run(self, plan, x, final_planar=None):
        """``final_planar``: the decoder head's output tensor (planar, clamped), or "alloc" to allocate it here."""
        pre = None                 # silu(norm(x)) handed over by the previous residual block's epilogue
        for i, (kind, name, cin, cout) in enumerate(plan):
            if kind == "conv":
                x, pre = self._conv(x, name, TAPS_333, cout), None
            elif kind == "res":
                nxt = plan[i + 1] if i + 1 < len(plan) else None
                ng = None
                if nxt is not None and cout <= self.FUSE_NORM_MAX_C and nxt[0] in ("res", "head"):
                    ng = self.w[nxt[1] + (".residual.0.gamma" if nxt[0] == "res" else ".0.gamma")]
                x, pre = self._res(x, name, cin, cout, pre=pre, next_gamma=ng)
            elif kind == "attn":
                x, pre = self._attn(x, name, cin), None
            elif kind in ("down2d", "down3d"):
                x, pre = self._down(x, name, cin, kind == "down3d"), None
            elif kind in ("up2d", "up3d"):
                x, pre = self._up(x, name, cin, kind == "up3d"), None
            elif kind == "head":
                y = pre if pre is not None else lib.rms_norm_cl(x, self.w[name + ".0.gamma"])
                pre = None
                if final_planar is not None:
                    T, H, Wd, _ = y.shape
                    if isinstance(final_planar, str):
                        final_planar = torch.empty(cout, T, H, Wd, dtype=F32, device=y.device)
                    lib.conv_tf32(y, self.w[name + ".2.w"], self.w[name + ".2.b"], TAPS_333, final_planar, T=T, H=H, W=Wd,
                                  Cout=cout, planar_clamp=True, tile_w=self._tile_w(Wd))
                    x = final_planar
                else:
                    x = self._conv(y, name + ".2", TAPS_333, cout, round_out=True)   # encoder head -> the 1x1 conv1 only
        return x


    # ------------------------------------------------------------------------------ row sharding
    UPS, DOWNS = ("up2d", "up3d"), ("down2d", "down3d")

    def enable_row_sharding(self, group=None, world: int = None, rank: int = None):
        """Split encode / decode across the ranks of ``group`` by image rows (see the module docstring)."""
        import torch.distributed as dist
        self.shard = SimpleNamespace(group=group, world=dist.get_world_size(group) if world is None else world,
                                     rank=dist.get_rank(group) if rank is None else rank)
        return self.shard

    @staticmethod
    def row_bounds(h: int, world: int) -> List[int]:
        """Latent-resolution row ranges of the ranks: rank r owns [b[r], b[r+1])."""
        return [(r * h) // world for r in range(world + 1)]

    @classmethod
    def _needed_rows(cls, seg, out_range: Tuple[int, int], h_in: int):
        """Walk ``seg`` backwards: need[i] = rows of layer i's INPUT (clamped to the image) that rows ``out_range`` of
        the segment's output depend on; need[len(seg)] = out_range.  Also returns the image height at every layer."""
        hs = [h_in]
        for kind, *_ in seg:
            hs.append(hs[-1] * 2 if kind in cls.UPS else hs[-1] // 2 if kind in cls.DOWNS else hs[-1])
        need = [None] * (len(seg) + 1)
        need[-1] = (max(out_range[0], 0), min(out_range[1], hs[-1]))
        for i in range(len(seg) - 1, -1, -1):
            lo, hi = need[i + 1]
            kind = seg[i][0]
            if kind in ("conv", "head"):
                lo, hi = lo - 1, hi + 1                       # one 3x3(x3) convolution
            elif kind == "res":
                lo, hi = lo - 2, hi + 2                       # two 3x3x3 convolutions (the shortcut is 1x1)
            elif kind in cls.UPS:
                lo, hi = lo // 2 - 1, (hi + 1) // 2 + 1       # out row 2y+p reads rows y-1..y (p=0) / y..y+1 (p=1)
            elif kind in cls.DOWNS:
                lo, hi = 2 * lo, 2 * hi + 2                   # out row y reads rows 2y..2y+2 (+ the zero pad row)
            else:
                raise ValueError(f"layer kind {kind!r} is not local in space")
            need[i] = (max(lo, 0), min(hi, hs[i]))
        return need, hs

    def _run_rows(self, seg, x, a: int, need, final_planar_rows=None):
        """Run ``seg`` on a row slab: ``x`` [T, rows, W, C] holds image rows a.. of the segment's input.  Before every
        resampling layer the slab is cut down to the rows still needed; the layers between two cuts run as ONE ``_run`` call,
        so the cross-layer epilogue fusions are the same as in the unsharded evaluation (bit-identical results).
        Returns (output slab, its first image row)."""
        i = 0
        while i < len(seg):
            j = i + 1
            while j < len(seg) and seg[j][0] not in self.UPS and seg[j][0] not in self.DOWNS:
                j += 1
            lo, hi = need[i]
            if lo > a or hi < a + x.shape[1]:
                x = x[:, lo - a:hi - a].contiguous()
                a = lo
            kind = seg[i][0]
            last_is_planar_head = seg[j - 1][0] == "head" and final_planar_rows is not None
            x = self._run(seg[i:j], x, final_planar="alloc" if last_is_planar_head else None)
            a = a * 2 if kind in self.UPS else a // 2 if kind in self.DOWNS else a
            if last_is_planar_head:
                lo, hi = final_planar_rows
                return x[:, :, lo - a:hi - a], lo
            i = j
        lo, hi = need[-1]
        return x[:, lo - a:hi - a], lo

    # Where a row stage ends besides the attention: after every downsampling layer of the encoder and before the first two
    # upsampling layers of the decoder - i.e. at the SMALLEST tensor between two resolution levels.  Inside one stage a
    # rank's halo grows by a row per 3x3 convolution and doubles at every resampling layer towards the high-resolution
    # side, so a stage that spans all four levels recomputes 3.8x (encode) / 2.2x (decode) of its share at 8 ranks; cut per
    # level the halo restarts at ~5-7 rows of the level's own resolution (1.2x at full resolution), for an all-gather of
    # 0.2-3 GB of features per cut.  The last upsampling layer is not cut (its input is 6 GB at 81 x 240 x 416 x 192).
    level_cuts = True

    def _segments(self, plan):
        """Cut a layer plan into row / frame stages: [("rows", layers), ("frames", [attn]), ("rows", layers), ...].
        Everything but the mid-block attention is local in space (row-shardable with a halo); the attention is one
        full-frame softmax per frame and shards by frames instead.  Cuts fall only where the unsharded evaluation fuses
        nothing across (before / after a resampling layer, around the attention), so the stitched result stays bit-identical."""
        ups_left = sum(1 for l in plan if l[0] in self.UPS)
        segs, cur = [], []
        for layer in plan:
            if layer[0] == "attn":
                if cur:
                    segs.append(("rows", cur))
                segs.append(("frames", [layer]))
                cur = []
                continue
            if self.level_cuts and layer[0] in self.UPS:
                if ups_left > 1 and cur:
                    segs.append(("rows", cur))
                    cur = []
                ups_left -= 1
            cur.append(layer)
            if self.level_cuts and layer[0] in self.DOWNS:
                segs.append(("rows", cur))
                cur = []
        if cur:
            segs.append(("rows", cur))
        return segs

    def sharded_stages(self, which: str, x: torch.Tensor, world: int):
        """The sharded evaluation of encode (``which='enc'``, x = [3,F,H,W] planar video) or decode (``'dec'``, x =
        [T,h,w,z] channels-last output of conv2) as a list of stages.  A stage is ``fn(rank, full) -> (local, dim, bounds)``:
        rank's share of the stage output computed from the stage's full (replicated) input, to be gathered along ``dim``
        with ``bounds`` before the next stage.  The last stage's gathered output is the result (encode: channels-last
        [T',h,w,2z] after conv1; decode: planar [3,F,H,W]).  ``encode`` / ``decode`` run the stages with an all-gather
        between them; the parity test runs every rank's share on one GPU and stitches."""
        sc = self.spatial_scale
        segs = self._segments(self.enc_plan if which == "enc" else self.dec_plan)
        stages = []
        if which == "enc":
            h_lat = x.shape[2] // sc
        else:
            h_lat = x.shape[1]
        rb = self.row_bounds(h_lat, world)
        for si, (kind, layers) in enumerate(segs):
            first, last = si == 0, si == len(segs) - 1
            if kind == "frames":
                def attn_stage(rank, full, layers=layers):
                    fb = self.row_bounds(full.shape[0], world)          # frames split like rows: as evenly as possible
                    kind_, name, cin, _ = layers[0]
                    part = self._attn(full[fb[rank]:fb[rank + 1]].contiguous(), name, cin) if fb[rank + 1] > fb[rank] \
                        else full[:0]
                    return part, 0, fb
                stages.append(attn_stage)
                continue
            # a row segment: its output resolution relative to the latent grid decides the row bounds
            n_up = sum(1 for l in layers if l[0] in self.UPS)
            n_down = sum(1 for l in layers if l[0] in self.DOWNS)

            def row_stage(rank, full, layers=layers, first=first, last=last, n_up=n_up, n_down=n_down):
                planar_in = which == "enc" and first                   # the video arrives planar [3,F,H,W]
                h_in = full.shape[2] if planar_in else full.shape[1]
                h_out = (h_in << n_up) >> n_down
                f = h_out // h_lat                                     # output rows per latent row (1 or spatial_scale)
                bounds = [f * v for v in rb]
                lo, hi = bounds[rank], bounds[rank + 1]
                need, _ = self._needed_rows(layers, (lo, hi), h_in)
                a, b = need[0]
                if planar_in:
                    slab = lib.planar_to_cl(full[:, :, a:b].to(F32).contiguous(), 4, round_tf32=True)
                else:
                    slab = full[:, a:b].contiguous()
                if which == "dec" and last:                            # the head writes the clamped planar clip
                    y, _ = self._run_rows(layers, slab, a, need, final_planar_rows=(lo, hi))
                    return y, 2, bounds
                y, _ = self._run_rows(layers, slab, a, need)
                if which == "enc" and last:                            # conv1 (1x1) on the rank's rows
                    y = self._conv(y.contiguous(), "conv1", TAPS_1, 2 * self.z_dim)
                return y, 1, bounds
            stages.append(row_stage)
        return stages

    def _all_gather_rows(self, local: torch.Tensor, dim: int, bounds: List[int]) -> torch.Tensor:
        """``local`` holds rows [bounds[r], bounds[r+1]) along ``dim``; returns the tensor with all rows, same on every rank."""
        import torch.distributed as dist
        sh = self.shard
        mx = max(bounds[r + 1] - bounds[r] for r in range(sh.world))
        shape = list(local.shape); shape[dim] = mx
        send = torch.zeros(shape, dtype=local.dtype, device=local.device)
        send.narrow(dim, 0, local.shape[dim]).copy_(local)
        recv = torch.empty([sh.world] + shape, dtype=local.dtype, device=local.device)
        dist.all_gather(list(recv.unbind(0)), send, group=sh.group)
        return assemble_rows([recv[r].narrow(dim, 0, bounds[r + 1] - bounds[r]) for r in range(sh.world)], dim)

    # ------------------------------------------------------------------------------ public surface
    @torch.no_grad()
    def encode(self, x: torch.Tensor):
        """x [1,3,F,H,W] fp32 in [-1,1] -> object with ``.latent_dist.mode()`` = mean latents [1,z,f,H/8,W/8]."""
        if not x.is_cuda:
            raise lib.WfError("WfWanVAE runs on CUDA tensors only (no CPU fallback)")
        assert x.shape[0] == 1 and x.shape[1] == 3
        if self.shard is not None and self.shard.world > 1:
            h = x[0]
            for stage in self.sharded_stages("enc", h, self.shard.world):
                part, dim, bounds = stage(self.shard.rank, h)
                h = self._all_gather_rows(part, dim, bounds)
        else:
            cl = lib.planar_to_cl(x[0].to(F32).contiguous(), 4, round_tf32=True)
            h = self._run(self.enc_plan, cl)
            h = self._conv(h, "conv1", TAPS_1, 2 * self.z_dim)
        mu = lib.cl_to_planar(h, self.z_dim)
        return SimpleNamespace(latent_dist=_Dist(mu.unsqueeze(0)))

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = False):
        """z [1,z_dim,f,h,w] (un-normalised latents) -> ([1,3,4(f-1)+1,8h,8w] fp32 clamped to [-1,1],)"""
        if not z.is_cuda:
            raise lib.WfError("WfWanVAE runs on CUDA tensors only (no CPU fallback)")
        assert z.shape[0] == 1 and z.shape[1] == self.z_dim
        f, h, w = z.shape[2:]
        cl = lib.planar_to_cl(z[0].to(F32).contiguous(), self.z_dim, round_tf32=True)
        x = self._conv(cl, "conv2", TAPS_1, self.z_dim, round_out=True)        # conv2 -> decoder.conv1 only
        nt = 0
        for i, up in enumerate(self.temperal_downsample):
            nt += 1 if up else 0
        F_out = (f - 1) * (2 ** nt) + 1
        if self.shard is not None and self.shard.world > 1:
            for stage in self.sharded_stages("dec", x, self.shard.world):
                part, dim, bounds = stage(self.shard.rank, x)
                x = self._all_gather_rows(part, dim, bounds)
            return (x.unsqueeze(0),)
        out = torch.empty(3, F_out, 8 * h, 8 * w, dtype=F32, device=z.device)
        self._run(self.dec_plan, x, final_planar=out)
        return (out.unsqueeze(0),)

    @classmethod
    def random_init(cls, device, seed: int = 4321, **kw):
        """Random-init weights with the reference network's shapes."""
        return cls(random_state_dict(seed=seed, **kw), device, **kw)


def random_state_dict(seed: int = 4321, **kw) -> Dict[str, torch.Tensor]:
    """A random state dict under the reference's ``WanVAE_`` parameter names (CPU tensors)."""
    proto = WfWanVAE.__new__(WfWanVAE)
    proto.dim = kw.get("dim", 96); proto.z_dim = kw.get("z_dim", 16)
    proto.dim_mult = tuple(kw.get("dim_mult", (1, 2, 4, 4))); proto.num_res_blocks = kw.get("num_res_blocks", 2)
    proto.temperal_downsample = list(kw.get("temporal_downsample", (False, True, True)))
    enc, dec = proto._plans()
    g = torch.Generator().manual_seed(seed)
    sd = {}
    def conv(name, cin, cout, k):
        fan = cin * k[0] * k[1] * k[2] if len(k) == 3 else cin * k[0] * k[1]
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) / fan ** 0.5
        sd[name + ".bias"] = 0.02 * torch.randn(cout, generator=g)
    def gam(name, c, nd):
        sd[name] = (1.0 + 0.05 * torch.randn(c, generator=g)).reshape(c, *([1] * nd))
    for plan in (enc, dec):
        for kind, name, cin, cout in plan:
            if kind == "conv":
                conv(name, cin, cout, (3, 3, 3))
            elif kind == "res":
                gam(name + ".residual.0.gamma", cin, 3); conv(name + ".residual.2", cin, cout, (3, 3, 3))
                gam(name + ".residual.3.gamma", cout, 3); conv(name + ".residual.6", cout, cout, (3, 3, 3))
                if cin != cout:
                    conv(name + ".shortcut", cin, cout, (1, 1, 1))
            elif kind == "attn":
                gam(name + ".norm.gamma", cin, 2); conv(name + ".to_qkv", cin, 3 * cin, (1, 1)); conv(name + ".proj", cin, cin, (1, 1))
            elif kind in ("down2d", "down3d"):
                conv(name + ".resample.1", cin, cin, (3, 3))
                if kind == "down3d":
                    conv(name + ".time_conv", cin, cin, (3, 1, 1))
            elif kind in ("up2d", "up3d"):
                conv(name + ".resample.1", cin, cin // 2, (3, 3))
                if kind == "up3d":
                    conv(name + ".time_conv", cin, 2 * cin, (3, 1, 1))
            elif kind == "head":
                gam(name + ".0.gamma", cin, 3); conv(name + ".2", cin, cout, (3, 3, 3))
    conv("conv1", 2 * proto.z_dim, 2 * proto.z_dim, (1, 1, 1)); conv("conv2", proto.z_dim, proto.z_dim, (1, 1, 1))
    return sd

