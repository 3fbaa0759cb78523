"""LongCat block-sparse attention on the GPU vs the oracle (oracle/longcat_bsa.py, pinned to the reference's Triton kernel).

Reference: longcat_video/block_sparse_attention/bsa_interface.py:612-659, flash_attn_bsa_varlen_mask.py:174-285.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _qkv(grid_q, grid_k, heads, seed, device):
    g = torch.Generator().manual_seed(seed)
    Lq, Lk = math.prod(grid_q), math.prod(grid_k)
    q = torch.randn(Lq, heads * 128, generator=g).to(BF)
    k = torch.randn(Lk, heads * 128, generator=g).to(BF)
    v = torch.randn(Lk, heads * 128, generator=g).to(BF)
    return q, k, v


def _bhsd(x, heads):
    return x.reshape(x.shape[0], heads, 128).permute(1, 0, 2).unsqueeze(0)


def _oracle_sparse(q, k, v, idx, lens, grid_q, grid_k, chunk, heads):
    from oracle import longcat_bsa as ob
    c = math.prod(chunk)
    qb, kb, vb = (ob.to_blocks(_bhsd(t, heads), g, chunk) for t, g in ((q, grid_q), (k, grid_k), (v, grid_k)))
    o = ob.sparse_attention(qb, kb, vb, idx.unsqueeze(0), lens.unsqueeze(0), c, c, 128 ** -0.5)
    return ob.from_blocks(o, grid_q, chunk)[0].permute(1, 0, 2).reshape(q.shape[0], heads * 128)


@pytest.mark.parametrize("chunk", [(4, 4, 4), (4, 4, 8)])
def test_mean_pool(cuda, chunk):
    from oracle import longcat_bsa as ob
    from worldforge_b200 import lib
    grid, heads = (8, 8, 16), 3
    q, _, _ = _qkv(grid, grid, heads, 0, cuda)
    got = lib.bsa_mean_pool(q.to(cuda), grid, chunk, heads).cpu()
    exp = ob.mean_pool(ob.to_blocks(_bhsd(q, heads), grid, chunk), math.prod(chunk))[0]
    # fp32 sums in a different order, then one bf16 rounding: equal up to rare 1-ulp flips
    d = (got.float() - exp.float()).abs()
    assert (d > 0).float().mean() < 5e-3 and d.max() <= 2 ** -7 * exp.float().abs().max()


@pytest.mark.parametrize("Nq,Nk,n_sel", [(16, 16, 8), (37, 70, 9), (9, 300, 37), (5, 33, 33), (8, 64, 1)])
def test_select_topk(cuda, Nq, Nk, n_sel):
    from worldforge_b200 import lib
    heads = 3
    g = torch.Generator().manual_seed(Nq * 1000 + Nk)
    qc = torch.randn(heads, Nq, 128, generator=g).to(BF).to(cuda)
    kc = torch.randn(heads, Nk, 128, generator=g).to(BF).to(cuda)
    idx = lib.bsa_select_topk(qc, kc, n_sel)
    assert idx.shape == (heads, Nq, n_sel)
    score = torch.matmul(qc, kc.transpose(-1, -2))            # bf16 x bf16 -> bf16, as cal_score does on the GPU
    assert (idx[..., 1:] > idx[..., :-1]).all()                # ascending, distinct
    assert int(idx.min()) >= 0 and int(idx.max()) < Nk
    ref = torch.topk(score, n_sel)[1]
    thr = torch.gather(score, -1, ref).float().min(dim=-1, keepdim=True).values     # the n_sel-th largest score
    got_scores = torch.gather(score, -1, idx.long()).float()
    # every chosen chunk scores at least the threshold, up to one bf16 ulp of accumulation-order noise
    ulp = thr.abs() * 2 ** -7 + 1e-6
    assert (got_scores >= thr - ulp).all()
    # and nothing clearly above the threshold was left out
    chosen = torch.zeros_like(score, dtype=torch.bool).scatter_(-1, idx.long(), True)
    assert not ((score.float() > thr + ulp) & ~chosen).any()


def test_select_topk_ties(cuda):
    """All scores equal: the lowest chunk indices win."""
    from worldforge_b200 import lib
    qc = torch.zeros(2, 4, 128, dtype=BF, device=cuda)
    kc = torch.randn(2, 50, 128, device=cuda).to(BF)
    idx = lib.bsa_select_topk(qc, kc, 7)
    assert (idx.cpu() == torch.arange(7, dtype=torch.int32)).all()


CASES = [
    # grid_q, grid_k, chunk, heads, n_sel
    ((4, 8, 8), (4, 8, 8), (4, 4, 4), 2, 2),          # 4 query chunks of 64: one CTA, both tiles full
    ((4, 4, 12), (8, 4, 12), (4, 4, 4), 1, 3),        # 3 query chunks (odd): half-empty tile; Tq != Tk (noise rows vs all keys)
    ((8, 8, 8), (8, 8, 8), (4, 4, 8), 2, 3),          # chunks of 128
    ((4, 4, 8), (12, 4, 8), (4, 4, 8), 1, 2),         # one 128-chunk: second tile empty
    ((8, 12, 16), (8, 12, 16), (4, 4, 4), 3, 11),     # several CTAs, long lists (more than one 16-entry window)
]


@pytest.mark.parametrize("grid_q,grid_k,chunk,heads,n_sel", CASES)
@pytest.mark.parametrize("varlen", [False, True])
def test_sparse_attention(cuda, grid_q, grid_k, chunk, heads, n_sel, varlen):
    from worldforge_b200 import lib
    c = math.prod(chunk)
    nq, nk = math.prod(grid_q) // c, math.prod(grid_k) // c
    q, k, v = _qkv(grid_q, grid_k, heads, 3, cuda)
    g = torch.Generator().manual_seed(11)
    idx = torch.stack([torch.stack([torch.randperm(nk, generator=g)[:n_sel] for _ in range(nq)]) for _ in range(heads)]).to(torch.int32)
    if varlen:
        lens = torch.randint(0, n_sel + 1, (heads, nq), generator=g, dtype=torch.int32)
        lens[0, 0] = 0                                            # an empty selection gives zeros
        lens[-1, -1] = n_sel
    else:
        lens = torch.full((heads, nq), n_sel, dtype=torch.int32)
    exp = _oracle_sparse(q, k, v, idx, lens, grid_q, grid_k, chunk, heads)
    out = torch.full_like(q, float("nan")).to(cuda)
    lib.attention_bsa_bf16(q.to(cuda), k.to(cuda), v.to(cuda), out, heads, idx.to(cuda), lens.to(cuda) if varlen else None,
                           grid_q, grid_k, chunk)
    got = out.cpu().float()
    assert torch.isfinite(got).all()
    err = (got - exp.float()).abs().max().item()
    assert err < 2e-2, err                     # bf16 outputs of magnitude ~1; the dense kernel's tolerance
    rel = ((got - exp.float()).norm() / exp.float().norm()).item()
    assert rel < 6e-3, rel


def test_dense_equivalence(cuda):
    """Selecting every key chunk reproduces the dense attention kernel's result."""
    from worldforge_b200 import lib
    grid, chunk, heads = (4, 8, 8), (4, 4, 4), 2
    q, k, v = (t.to(cuda) for t in _qkv(grid, grid, heads, 5, cuda))
    nq = math.prod(grid) // 64
    idx = torch.arange(nq, dtype=torch.int32, device=cuda).expand(heads, nq, nq).contiguous()
    a, b = torch.empty_like(q), torch.empty_like(q)
    lib.attention_bsa_bf16(q, k, v, a, heads, idx, None, grid, grid, chunk)
    lib.attention_bf16(q, k, v, b, heads)
    assert (a.float() - b.float()).abs().max().item() < 2e-2


@pytest.mark.parametrize("Nq,Nk,thr,sparsity", [(16, 16, 0.5, None), (37, 70, 0.9, None), (9, 300, 0.6, 0.9), (5, 33, 0.99, 0.5),
                                                (8, 880, 0.8, 0.9375)])
def test_select_cdf(cuda, Nq, Nk, thr, sparsity):
    """get_select_indices_cdf(_topk) (bsa_interface.py:234-275) on the device against the oracle's torch restatement: the
    sorted order is the descending order of the bf16 scores, and the selected count equals the oracle's up to the one chunk
    whose cumulative mass sits within fp32 summation noise of the threshold."""
    from oracle import longcat_bsa as ob
    from worldforge_b200 import lib
    heads = 3
    g = torch.Generator().manual_seed(Nq * 1000 + Nk)
    qc = (torch.randn(heads, Nq, 128, generator=g) * 1.5).to(BF).to(cuda)
    kc = torch.randn(heads, Nk, 128, generator=g).to(BF).to(cuda)
    n_floor = int((1 - sparsity) * Nk) if sparsity is not None else 0
    idx, lens = lib.bsa_select_cdf(qc, kc, thr, n_floor)
    assert idx.shape == (heads, Nq, Nk) and lens.shape == (heads, Nq)
    score = torch.matmul(qc, kc.transpose(-1, -2)).float().cpu()            # bf16 x bf16 -> bf16, as cal_score does on the GPU
    idx_c, lens_c = idx.cpu().long(), lens.cpu()
    assert (torch.sort(idx_c, dim=-1).values == torch.arange(Nk)).all()     # a permutation of the key chunks
    sorted_scores = torch.gather(score, -1, idx_c)
    ulp = sorted_scores.abs() * 2 ** -7 + 1e-6
    assert (sorted_scores[..., :-1] + ulp[..., :-1] >= sorted_scores[..., 1:]).all()     # descending up to accumulation-order noise
    _, want = ob.select(score.unsqueeze(0), sparsity, thr, 128)
    want = want[0]
    assert (lens_c - want).abs().max() <= 1, (lens_c - want).abs().max()
    assert ((lens_c - want) != 0).float().mean() < 0.1
    assert int(lens_c.min()) >= n_floor and int(lens_c.max()) <= Nk


def test_cdf_selection_through_the_attention(cuda):
    """enable_bsa() with a cdf_threshold: the selected lists (variable lengths) drive the sparse kernel like the oracle's."""
    from worldforge_b200 import lib
    grid, chunk, heads = (8, 8, 16), (4, 4, 4), 2
    q, k, v = _qkv(grid, grid, heads, 3, cuda)
    qd, kd, vd = q.to(cuda), k.to(cuda), v.to(cuda)
    q_cmp, k_cmp = lib.bsa_mean_pool(qd, grid, chunk, heads), lib.bsa_mean_pool(kd, grid, chunk, heads)
    idx, lens = lib.bsa_select_cdf(q_cmp, k_cmp, 0.7, 2)
    out = torch.empty_like(qd)
    lib.attention_bsa_bf16(qd, kd, vd, out, heads, idx, lens, grid, grid, chunk)
    want = _oracle_sparse(q, k, v, idx.cpu(), lens.cpu(), grid, grid, chunk, heads)
    rel = ((out.cpu().float() - want.float()).norm() / want.float().norm()).item()
    assert rel < 8e-3, rel
    assert 2 <= int(lens.min()) and int(lens.max()) < 16          # a proper subset of the 16 key chunks (sorted lists, lens from the mass rule)
