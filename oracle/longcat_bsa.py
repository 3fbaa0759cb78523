"""CPU restatement of LongCat-Video's block-sparse attention (the 720p refine pass), plain PyTorch.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Reference: longcat_for_worldforge/longcat_video/
block_sparse_attention/bsa_interface.py (gating + wrapper) and flash_attn_bsa_varlen_mask.py:174-285 (the Triton
forward kernel).  Pinned against both, the Triton kernel run by Triton's CPU interpreter (oracle/make_golden.py,
tests/test_oracle_pinning.py).

  to_blocks / from_blocks    [T,H,W] token order <-> chunk-major order            bsa_interface.py:598-610
  mean_pool                  per-chunk mean of q / k in the tensor's dtype         :176-186
  select_topk                score = q_cmp k_cmp^T, top int((1-sparsity) Nk)       :188-192,214-232
  select_cdf / cdf_topk      softmax mass threshold (optionally floored by top-k)  :234-275
  sparse_attention           softmax over the selected key chunks only             flash_attn_bsa_varlen_mask.py:236-284
  bsa_3d                     the whole of flash_attn_bsa_3d                        bsa_interface.py:612-659
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch


def to_blocks(x: torch.Tensor, grid: Sequence[int], chunk: Sequence[int]) -> torch.Tensor:
    """[B,H,T*Hh*W,D] in (t,h,w) order -> chunk-major order (Nt,Nh,Nw,t,h,w)."""
    B, Hn, _, D = x.shape
    (T, Hh, W), (t, h, w) = grid, chunk
    x = x.reshape(B, Hn, T // t, t, Hh // h, h, W // w, w, D).permute(0, 1, 2, 4, 6, 3, 5, 7, 8)
    return x.reshape(B, Hn, T * Hh * W, D)


def from_blocks(x: torch.Tensor, grid: Sequence[int], chunk: Sequence[int]) -> torch.Tensor:
    B, Hn, _, D = x.shape
    (T, Hh, W), (t, h, w) = grid, chunk
    x = x.reshape(B, Hn, T // t, Hh // h, W // w, t, h, w, D).permute(0, 1, 2, 5, 3, 6, 4, 7, 8)
    return x.reshape(B, Hn, T * Hh * W, D)


def block_permutation(grid: Sequence[int], chunk: Sequence[int]) -> torch.Tensor:
    """perm[i] = the (t,h,w)-order token index that sits at position i of the chunk-major order."""
    T, Hh, W = grid
    idx = torch.arange(T * Hh * W).reshape(1, 1, -1, 1)
    return to_blocks(idx, grid, chunk).reshape(-1)


def mean_pool(x: torch.Tensor, block: int) -> torch.Tensor:
    B, Hn, S, D = x.shape
    nb = math.ceil(S / block)
    if S % block:
        x = torch.nn.functional.pad(x, (0, 0, 0, nb * block - S))
    return x.reshape(B, Hn, nb, block, D).mean(dim=3)


def scores(q_cmp: torch.Tensor, k_cmp: torch.Tensor) -> torch.Tensor:
    return torch.matmul(q_cmp, k_cmp.transpose(-1, -2))


def select(score: torch.Tensor, sparsity: Optional[float], cdf_threshold: Optional[float], head_dim: int):
    """-> (indices [B,H,Nq,S], lens [B,H,Nq]); only indices[..., :lens] are attended."""
    nk = score.shape[-1]
    if sparsity is not None and cdf_threshold is None:
        n = int((1 - sparsity) * nk)
        idx = torch.topk(score, n)[1]
        return idx, torch.full(idx.shape[:3], n, dtype=torch.int32)
    w = torch.softmax(score * (1 / head_dim ** 0.5), dim=-1)
    ws = torch.sort(w, dim=-1, descending=True)
    cdf = torch.cumsum(ws.values, dim=-1)
    thr = torch.full(cdf.shape[:3] + (1,), cdf_threshold, dtype=cdf.dtype)
    n = torch.searchsorted(cdf, thr, right=True)
    if sparsity is not None:
        n = n.clamp(min=int((1 - sparsity) * nk))
    return ws.indices, n.squeeze(-1)


def sparse_attention(q, k, v, idx, lens, chunk_q: int, chunk_k: int, scale: float) -> torch.Tensor:
    """Per q-chunk: softmax(q k_sel^T * scale) v_sel over its selected key chunks; fp32 scores and statistics, the
    probabilities rounded to v's dtype for the PV product, fp32 accumulation, output in q's dtype."""
    B, Hn, S, D = q.shape
    out = torch.empty_like(q)
    for b in range(B):
        for h in range(Hn):
            for m in range(S // chunk_q):
                sel = idx[b, h, m, : int(lens[b, h, m])].to(torch.int64)
                if sel.numel() == 0:                       # the kernel's initial state: acc = 0, l = 1 (:230-232)
                    out[b, h, m * chunk_q:(m + 1) * chunk_q] = 0
                    continue
                keys = (sel[:, None] * chunk_k + torch.arange(chunk_k)[None]).reshape(-1)
                qq = q[b, h, m * chunk_q:(m + 1) * chunk_q].float()
                s = qq @ k[b, h, keys].float().T * scale
                mx = s.max(dim=-1, keepdim=True).values
                p = torch.exp(s - mx)
                l = p.sum(-1, keepdim=True)
                o = p.to(v.dtype).float() @ v[b, h, keys].float() / l
                out[b, h, m * chunk_q:(m + 1) * chunk_q] = o.to(q.dtype)
    return out


def bsa_3d(q, k, v, grid_q, grid_k, sparsity=0.875, cdf_threshold=None, chunk_q=(4, 4, 8), chunk_k=(4, 4, 8),
           return_selection: bool = False):
    D = q.shape[-1]
    qb, kb, vb = to_blocks(q, grid_q, chunk_q), to_blocks(k, grid_k, chunk_k), to_blocks(v, grid_k, chunk_k)
    cq, ck = math.prod(chunk_q), math.prod(chunk_k)
    sc = scores(mean_pool(qb, cq), mean_pool(kb, ck))
    idx, lens = select(sc, sparsity, cdf_threshold, D)
    o = from_blocks(sparse_attention(qb, kb, vb, idx, lens, cq, ck, 1 / D ** 0.5), grid_q, chunk_q)
    return (o, idx, lens, sc) if return_selection else o
