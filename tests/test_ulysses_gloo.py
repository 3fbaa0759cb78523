"""Sequence-parallel (Ulysses) host logic on 2 and 4 CPU ranks over gloo: the head<->token all-to-all layouts reproduce
single-rank attention, and the sharded head scatter + sum reproduces the gather.

Rendezvous is a file store (no TCP port to race for: a port probed free can be taken again before gloo binds it, which
made this test flaky in round 1), and the bit-for-bit comparison evaluates the single-rank reference with the SAME call
shapes (one call per rank's head group, one thread) as the ranks do - a batched CPU GEMM of another shape or thread count
may sum in another order, and one flipped bf16 rounding would fail ``torch.equal``."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wan_dit


def _store_file():
    fd, path = tempfile.mkstemp(prefix="wf_gloo_"); os.close(fd); os.unlink(path)
    return path


def _init(rank, world, store):
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{store}", rank=rank, world_size=world)


def _cpu_attn(q, k, v, out, heads):
    L = q.shape[0]
    o = wan_dit.attention(q.reshape(L, heads, 128), k.reshape(k.shape[0], heads, 128), v.reshape(v.shape[0], heads, 128), amp=True)
    out.copy_(o.reshape(L, heads * 128))


def _worker(rank, world, store, L, H, ret):
    _init(rank, world, store)
    try:
        from worldforge_b200 import ulysses
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(L, 3 * H * 128, generator=g).to(torch.bfloat16)
        Ll = L // world
        sp = ulysses.SequenceParallel()
        mine = sp.attention(qkv[rank * Ll:(rank + 1) * Ll].contiguous(), H, attn_fn=_cpu_attn)
        assert mine.shape == (Ll, H * 128) and sp.a2a_calls == 2
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        # the scatter-into-zero-canvas + all-reduce used for the head output
        canvas = torch.zeros(L, 4)
        canvas[rank * Ll:(rank + 1) * Ll] = rank + 1.0
        sp.all_reduce(canvas)
        if rank == 0:
            ret["out"] = torch.cat(parts).float()
            ret["canvas"] = canvas
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ulysses_attention_matches_single_rank(world):
    L, H = 64, 4
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _store_file(), L, H, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(L, 3 * H * 128, generator=g).to(torch.bfloat16)
    D, hp = H * 128, H // world * 128
    want = torch.empty(L, D, dtype=torch.bfloat16)
    old_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        for r in range(world):          # rank r attends over all tokens with heads [r*H/P, (r+1)*H/P): same call shape here
            sl = lambda j: qkv[:, j * D + r * hp: j * D + (r + 1) * hp].contiguous()
            part = torch.empty(L, hp, dtype=torch.bfloat16)
            _cpu_attn(sl(0), sl(1), sl(2), part, H // world)
            want[:, r * hp:(r + 1) * hp] = part
    finally:
        torch.set_num_threads(old_threads)
    assert torch.equal(ret["out"], want.float())
    c, Ll = ret["canvas"], L // world
    for r in range(world):
        assert (c[r * Ll:(r + 1) * Ll] == r + 1).all()


def test_layout_helpers_roundtrip():
    from worldforge_b200 import ulysses
    world, Ll, H = 4, 3, 8
    qkv = torch.arange(Ll * 3 * H * 128, dtype=torch.float32).view(Ll, 3 * H * 128)
    send = ulysses.heads_to_tokens_layout(qkv, world)
    hp = H // world * 128
    assert send.shape == (world, Ll, 3, hp)
    for d in range(world):          # block d carries heads [d*H/P, (d+1)*H/P) of q, k and v
        for j in range(3):
            assert torch.equal(send[d, :, j], qkv[:, j * H * 128 + d * hp: j * H * 128 + (d + 1) * hp])
    back = ulysses.tokens_to_heads_layout(torch.stack([qkv[:, :H * 128].view(Ll, world, hp)[:, r] for r in range(world)]))
    assert torch.equal(back, qkv[:, :H * 128])


def _peer_setup_worker(rank, world, store, ret):
    """PeerSequenceParallel's set-up must end the same way on every rank: here rank 1's allocation fails (there is no GPU
    at all in this test), and BOTH ranks must raise PeerSetupError instead of one of them waiting for the other."""
    _init(rank, world, store)
    try:
        from worldforge_b200 import lib, ulysses

        class FakeBuffer:
            def __init__(self, nbytes):
                if rank == 1:
                    raise lib.WfError("no device memory")
                self.ptr, self.handle, self.nbytes = 4096, b"h" * 64, nbytes
            open = staticmethod(lambda handle: 8192)
            freed = []

            def free(self):                                     # the failed set-up releases what this rank had allocated
                FakeBuffer.freed.append(self.nbytes)
            close_mapping = staticmethod(lambda ptr: None)

        lib.PeerBuffer = FakeBuffer
        try:
            ulysses.PeerSequenceParallel(None, 64, 32, 4, torch.device("cpu"))
            ret[rank] = "built"
        except ulysses.PeerSetupError as ex:
            ret[rank] = "agreed:" + str(ex) + f" freed={len(FakeBuffer.freed)}"
    finally:
        dist.destroy_process_group()


def test_peer_setup_failure_is_agreed_by_all_ranks():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_setup_worker, args=(2, _store_file(), ret), nprocs=2, join=True)
    assert ret[0].startswith("agreed:") and ret[1].startswith("agreed:")
    assert "no device memory" in ret[1] and "a peer could not" in ret[0]
    assert ret[0].endswith("freed=3") and ret[1].endswith("freed=0")      # rank 0 releases its three buffers, rank 1 had none


def _cfg_worker(rank, world, store, ret):
    _init(rank, world, store)
    try:
        from worldforge_b200 import ulysses
        sp_group, cfgp = ulysses.cfg_layout(world, rank)
        half = world // 2
        assert cfgp.branch == rank // half
        assert (sp_group is None) == (half == 1)
        if sp_group is not None:
            assert dist.get_world_size(sp_group) == half and dist.get_rank(sp_group) == rank % half
        # every rank of a half computes the same prediction (the Ulysses forward ends in an all-reduce inside the half)
        mine = torch.full((1, 4, 2, 3, 5), 10.0 * cfgp.branch + 1.0)
        v_c, v_u = cfgp.exchange(mine)
        ret[rank] = (v_c.unique().tolist(), v_u.unique().tolist(), cfgp.exchanges)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cfg_parallel_layout(world):
    """CFG x Ulysses layout: rank r and r + P/2 swap the conditional / unconditional predictions."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cfg_worker, args=(world, _store_file(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == ([1.0], [11.0], 1), (r, ret[r])


def _cp2d_worker(rank, world, store, split_hw, ret):
    _init(rank, world, store)
    try:
        from worldforge_b200 import ulysses
        g = torch.Generator().manual_seed(0)
        T, Hp, Wp, H = 3, 4, 6, 4
        qkv = torch.randn(T, Hp, Wp, 3 * H * 128, generator=g).to(torch.bfloat16)          # the whole clip's tokens on every rank
        loc = ulysses.split_2d(qkv, (1, 2), split_hw, rank).reshape(-1, 3 * H * 128)          # this rank's block, (T, H', W') order
        # split / gather round trip of a [C, T, H, W]-shaped tensor
        full = torch.arange(2 * T * Hp * Wp, dtype=torch.float32).view(2, T, Hp, Wp)
        parts = [torch.empty_like(ulysses.split_2d(full, (2, 3), split_hw, rank).contiguous()) for _ in range(world)]
        dist.all_gather(parts, ulysses.split_2d(full, (2, 3), split_hw, rank).contiguous())
        assert torch.equal(ulysses.gather_2d(parts, (2, 3), split_hw), full)
        D = H * 128
        sp = ulysses.GeneralSequenceParallel()
        nq = loc.shape[0] // 3                                                               # "noise" queries: the last two frames' tokens
        seen = {}
        def attn(q, k, v, out, heads):
            seen["shapes"] = (tuple(q.shape), tuple(k.shape), heads)
            _cpu_attn(q, k, v, out, heads)
        mine = sp.attention_qkv(loc[nq:, :D].contiguous(), loc[:, D:2 * D].contiguous(), loc[:, 2 * D:].contiguous(), H, attn)
        assert seen["shapes"] == ((world * (loc.shape[0] - nq), D // world), (world * loc.shape[0], D // world), H // world)
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine)
        if rank == 0:
            ret["outs"] = [o.float() for o in outs]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,split_hw", [(2, (1, 2)), (4, (2, 2))])
def test_context_parallel_2d_exchange(world, split_hw):
    """LongCat's 2-D context parallel (context_parallel_util.py:91-121, ulysses_wrapper.py:87-105) host logic: every rank
    keeps one block of every frame; q (a subset of the local tokens) / k / v are exchanged heads-for-tokens and the result
    equals single-rank attention of the same queries over ALL keys (dense attention does not depend on the key order)."""
    from worldforge_b200 import ulysses
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cp2d_worker, args=(world, _store_file(), split_hw, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    T, Hp, Wp, H = 3, 4, 6, 4
    D = H * 128
    qkv = torch.randn(T, Hp, Wp, 3 * D, generator=g).to(torch.bfloat16)
    flat = qkv.reshape(-1, 3 * D)
    for r in range(world):
        loc = ulysses.split_2d(qkv, (1, 2), split_hw, r).reshape(-1, 3 * D)
        nq = loc.shape[0] // 3
        want = torch.empty(loc.shape[0] - nq, D, dtype=torch.bfloat16)
        _cpu_attn(loc[nq:, :D].contiguous(), flat[:, D:2 * D].contiguous(), flat[:, 2 * D:].contiguous(), want, H)
        d = (ret["outs"][r] - want.float()).abs().max().item()
        assert d <= 2 ** -6, (r, d)                              # same math, keys in another order: bf16 rounding flips only
