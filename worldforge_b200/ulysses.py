"""Ulysses sequence parallelism for the DiT: tokens sharded across the GPUs of one box, heads exchanged around attention.

The clip's L tokens (frame-major, exactly the order WanModel flattens them, reference wan/modules/model.py:537) are cut
into P contiguous shards - the layout of the reference design wan/distributed/xdit_context_parallel.py:160-162 - and every
token-local op of the block (LayerNorm, the GEMMs, RMSNorm over the full width, RoPE with global positions,
cross-attention to the replicated 769-token context, FFN) runs on the local shard without communication.  Self-attention
needs every key: one all-to-all turns the fused q|k|v shard [L/P, 3, H, 128] into [L, 3, H/P, 128] (all tokens, this
rank's heads), the same attention kernel runs on H/P heads, and a second all-to-all returns [L/P, H, 128] - the
head<->sequence exchange of the reference's LongCat path (longcat_video/context_parallel/ulysses_wrapper.py:87-105), but
ONE collective for q, k and v instead of three and no second staging copy on the way in.  The 64-channel head output is
scattered by each rank into a zero canvas and summed (8 MB).  Weights are replicated (33 GB of 180 GB).

The scheduler and the per-pixel blend act on the replicated 8 MB latents and run identically on every rank; the VAE round
trip is split by image rows (frames for its mid-block attention) and the FLF channel scoring by channels.

``PeerSequenceParallel`` does the same exchange without collectives on the data path (NVLink peer stores from the producing
kernels); ``SequenceParallel`` is the NCCL all-to-all form it falls back to when peer mapping is unavailable.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def heads_to_tokens_layout(qkv: torch.Tensor, world: int) -> torch.Tensor:
    """[Ll, 3*H*128] (token shard, all heads) -> send buffer [world, Ll, 3*(H/world)*128]; block d goes to rank d."""
    Ll, W3 = qkv.shape
    hp = W3 // (3 * world)
    return qkv.view(Ll, 3, world, hp).permute(2, 0, 1, 3).contiguous()


def tokens_from_ranks(recv: torch.Tensor) -> torch.Tensor:
    """receive buffer [world(src), Ll, 3*hp] -> [L, 3*hp]: source rank r owns token block r, so this is already global order."""
    world, Ll = recv.shape[:2]
    return recv.reshape(world * Ll, -1)


def tokens_to_heads_layout(recv: torch.Tensor) -> torch.Tensor:
    """receive buffer of the return trip [world(src), Ll, hp] -> [Ll, world*hp] = [Ll, H*128] (heads of rank r at block r)."""
    world, Ll, hp = recv.shape
    return recv.permute(1, 0, 2).reshape(Ll, world * hp)


class SequenceParallel:
    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.a2a_calls = 0
        self.peer = False

    def all_to_all(self, send: torch.Tensor) -> torch.Tensor:
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        self.a2a_calls += 1
        return recv

    def all_reduce(self, t: torch.Tensor):
        dist.all_reduce(t, group=self.group)

    def attention(self, qkv: torch.Tensor, heads: int, attn_fn=None) -> torch.Tensor:
        """qkv [L/P, 3*heads*128] bf16 (RMSNorm + RoPE already applied) -> attention output [L/P, heads*128]."""
        from . import lib
        P = self.world
        hl = heads // P
        recv = self.all_to_all(heads_to_tokens_layout(qkv, P))          # [P, Ll, 3*hl*128]
        full = tokens_from_ranks(recv)                                   # [L, 3*hl*128]
        w = hl * 128
        out = torch.empty(full.shape[0], w, dtype=qkv.dtype, device=qkv.device)
        if attn_fn is None:
            lib.attention_bf16(full[:, :w], full[:, w:2 * w], full[:, 2 * w:], out, hl)
        else:
            attn_fn(full[:, :w], full[:, w:2 * w], full[:, 2 * w:], out, hl)
        back = self.all_to_all(out.view(P, full.shape[0] // P, w))       # block d = token shard d, already contiguous
        return tokens_to_heads_layout(back)


def split_2d(x: torch.Tensor, dim_hw, split_hw, rank: int) -> torch.Tensor:
    """split_tensor_in_cp_2d (longcat_video/context_parallel/context_parallel_util.py:91-121): the ``rank``-th block of ``x``
    cut into split_h x split_w blocks along dims ``dim_hw``, blocks numbered row-major (h outer, w inner)."""
    (dh, dw), (sh, sw) = dim_hw, split_hw
    if x.shape[dh] % sh or x.shape[dw] % sw:
        raise RuntimeError(f"sizes {x.shape[dh]} x {x.shape[dw]} are not multiples of the split {sh} x {sw}")
    bh, bw = x.shape[dh] // sh, x.shape[dw] // sw
    ih, iw = rank // sw, rank % sw
    return x.narrow(dh, ih * bh, bh).narrow(dw, iw * bw, bw)


def gather_2d(parts, dim_hw, split_hw) -> torch.Tensor:
    """Inverse of split_2d over the list of every rank's block (gather_cp_2d, context_parallel_util.py:180-235)."""
    (dh, dw), (sh, sw) = dim_hw, split_hw
    rows = [torch.cat(parts[ih * sw:(ih + 1) * sw], dim=dw) for ih in range(sh)]
    return torch.cat(rows, dim=dh)


class GeneralSequenceParallel(SequenceParallel):
    """Ulysses exchange for q, k, v of different lengths (LongCat: the noise tokens attend to all tokens, attention.py:124-135)
    and any local token set: rank r's tokens land at rows [r * Ll, (r+1) * Ll) of the gathered sequence - the rank-major
    order LongCat's context parallel produces and its block-sparse attention chunks in (attention.py:60-66)."""

    def _scatter_heads(self, x: torch.Tensor) -> torch.Tensor:
        """[Ll, H*128] (local tokens, all heads) -> [P*Ll, (H/P)*128] (all ranks' tokens rank-major, this rank's heads)."""
        P = self.world
        Ll, W = x.shape
        send = x.reshape(Ll, P, W // P).permute(1, 0, 2).contiguous()            # block d = heads of rank d
        return self.all_to_all(send).reshape(P * Ll, W // P)

    def _gather_heads(self, x: torch.Tensor) -> torch.Tensor:
        """[P*Ll, (H/P)*128] -> [Ll, H*128]."""
        P = self.world
        Ll = x.shape[0] // P
        back = self.all_to_all(x.reshape(P, Ll, x.shape[1]).contiguous())      # block d = token shard of rank d
        return back.permute(1, 0, 2).reshape(Ll, P * x.shape[1])

    def attention_qkv(self, q, k, v, heads: int, attn_fn) -> torch.Tensor:
        """q [Lq_l, H*128], k / v [Lk_l, H*128] (norm + RoPE applied) -> [Lq_l, H*128]; attn_fn(q, k, v, out, heads_local)."""
        hl = heads // self.world
        qf, kf, vf = self._scatter_heads(q), self._scatter_heads(k), self._scatter_heads(v)
        out = torch.empty_like(qf)
        attn_fn(qf, kf, vf, out, hl)
        return self._gather_heads(out)


class PeerSetupError(RuntimeError):
    """Raised on EVERY rank when any rank could not build the peer-memory exchange."""


class PeerSequenceParallel(SequenceParallel):
    """The same exchange without collectives on the data path: the kernels that produce q|k|v and the attention output
    store straight into the consuming rank's memory over NVLink (CUDA IPC peer mappings of ``wf_peer_alloc`` buffers).

      wf_qkv_norm_rope_scatter   RMSNorm + RoPE of q and k (and v) -> peer d's  full[L, 3*(H/P)*128]   (rows of this rank)
      barrier
      wf_attention_bf16_peers    attention over all tokens, this rank's heads -> row q to the rank owning token q,
                                 att[Ll, H*128] columns of this rank's heads
      barrier

    against NCCL's version of the reference pattern: two all-to-alls, three layout copies and two RMSNorm passes per
    layer.  The barrier is a 1-element all-reduce on the compute stream (stream-ordered after the producer kernel, so a
    rank's peers have finished writing when it returns).  ``att`` is double-buffered by layer parity: a rank may start
    writing layer l+1's output into a peer that is still reading layer l's (it cannot be two layers ahead - the barriers
    of layer l+1 need every rank's layer-l attention)."""

    def __init__(self, group, L: int, Ll: int, heads: int, device):
        super().__init__(group)
        from . import lib
        P, hl = self.world, heads // self.world
        self.L, self.Ll, self.heads, self.hl, self.device = L, Ll, heads, hl, device
        self.ld_full, self.ld_att = 3 * hl * 128, heads * 128
        # Set-up is collective and every step may fail on its own rank (allocation, IPC export, mapping a peer): the ranks
        # exchange what they have, try, and then AGREE on the outcome - either all of them use peer memory or all of them
        # raise PeerSetupError (the caller then keeps the NCCL all-to-all form; nothing hangs on a half-built exchange).
        err = None
        self._bufs, self._ptrs = [], []
        try:
            self._bufs = [lib.PeerBuffer(L * self.ld_full * 2), lib.PeerBuffer(Ll * self.ld_att * 2), lib.PeerBuffer(Ll * self.ld_att * 2)]
            mine = [b.handle for b in self._bufs]
        except Exception as ex:                       # noqa: BLE001 - reported to every rank below
            err, mine = ex, None
        handles = [None] * P
        dist.all_gather_object(handles, mine, group=self.group)
        if err is None and all(h is not None for h in handles):
            try:
                self._ptrs = [[self._bufs[i].ptr if r == self.rank else lib.PeerBuffer.open(handles[r][i]) for r in range(P)]
                              for i in range(3)]
            except Exception as ex:                   # noqa: BLE001
                err = ex
        elif err is None:
            err = RuntimeError("a peer could not allocate or export its buffers")
        oks = [None] * P
        dist.all_gather_object(oks, err is None, group=self.group)
        if not all(oks):
            self.close(sync=False)                      # whatever this rank allocated or mapped before the failure
            raise PeerSetupError(f"peer-memory exchange unavailable on rank(s) {[r for r, ok in enumerate(oks) if not ok]}: {err}")
        self.full = self._bufs[0].tensor((L, self.ld_full), torch.bfloat16, device)
        self.att = [self._bufs[1 + i].tensor((Ll, self.ld_att), torch.bfloat16, device) for i in range(2)]
        self._flag = torch.zeros(1, device=device)
        self.layer = 0
        self.barriers = 0

    def barrier(self):
        dist.all_reduce(self._flag, group=self.group)
        self.barriers += 1

    def close(self, sync: bool = True):
        """Unmap the peers' buffers and free this rank's (a change of resolution / frame count builds a new exchange; without
        this the old allocations and IPC mappings stayed for the life of the process).  Collective when ``sync``."""
        from . import lib
        if sync:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)              # no peer is still reading or writing these buffers
        for i, row in enumerate(self._ptrs):
            for r, p in enumerate(row):
                if r != self.rank:
                    lib.PeerBuffer.close_mapping(p)
        self._ptrs = []
        for b in self._bufs:
            b.free()
        self._bufs = []
        self.full, self.att = None, []

    def exchange_attention(self, qkv: torch.Tensor, norm_q, norm_k, rope, eps: float) -> torch.Tensor:
        """qkv [Ll, 3*H*128] bf16 straight from the projection (NOT yet normalised) -> attention output [Ll, H*128]."""
        from . import lib
        par = self.layer & 1
        self.layer += 1
        lib.qkv_norm_rope_scatter(qkv, norm_q, norm_k, rope, eps, self._ptrs[0], self.ld_full, self.rank * self.Ll)
        self.barrier()
        w = self.hl * 128
        col = self.rank * w * 2                                        # byte offset of this rank's head block in a row of att
        lib.attention_bf16_peers(self.full[:, :w], self.full[:, w:2 * w], self.full[:, 2 * w:],
                                 [p + col for p in self._ptrs[1 + par]], self.Ll, self.ld_att, self.hl)
        self.barrier()
        return self.att[par]


class CfgParallel:
    """The two forwards of classifier-free guidance (conditional / unconditional: pipeline_wan_i2v_clean.py:593-610) are
    independent, so on an even number of GPUs the ranks split into two halves that each run ONE of them - Ulysses inside a
    half, over half as many ranks (twice the token shard per rank, half the exchange volume, half the barriers; SURVEY.md
    §8e) - and rank r swaps its 8 MB prediction with rank r + P/2 before the CFG combination.  Both halves then hold both
    predictions and everything after the combination stays replicated exactly as before.  The sequence-parallel forward is
    bit-identical to the single-GPU one, so the result does not depend on the layout."""

    def __init__(self, pair_group, branch: int):
        self.group, self.branch = pair_group, branch           # branch 0: conditional, 1: unconditional
        self.exchanges = 0

    def exchange(self, v: torch.Tensor):
        """This rank's prediction -> (conditional, unconditional)."""
        both = [torch.empty_like(v), torch.empty_like(v)]
        dist.all_gather(both, v.contiguous(), group=self.group)
        self.exchanges += 1
        return both[0], both[1]


def cfg_layout(world: int, rank: int):
    """Process groups of the CFG x Ulysses layout: (sequence-parallel group of this rank's half or None, CfgParallel).
    Every rank creates every group, in the same order (torch.distributed requirement)."""
    if world < 2 or world % 2:
        raise ValueError("the CFG-parallel layout needs an even number of ranks")
    half = world // 2
    halves = [dist.new_group(list(range(b * half, (b + 1) * half))) if half > 1 else None for b in range(2)]
    pairs = [dist.new_group([r, r + half]) for r in range(half)]
    return halves[rank // half], CfgParallel(pairs[rank % half], rank // half)


def enable(transformer, group=None, peer: bool = False) -> SequenceParallel:
    """``peer=True``: exchange through NVLink peer memory (built lazily at the first forward, when the token count is known)."""
    sp = SequenceParallel(group)
    sp.peer = peer
    transformer.sp = sp
    return sp
