import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from worldforge_b200 import lib
import test_bsa_gpu as T
lib.load()
cuda = torch.device("cuda:0")
for case in T.CASES[:3]:
    grid_q, grid_k, chunk, heads, n_sel = case
    c = math.prod(chunk)
    nq, nk = math.prod(grid_q) // c, math.prod(grid_k) // c
    q, k, v = T._qkv(grid_q, grid_k, heads, 3, cuda)
    g = torch.Generator().manual_seed(11)
    idx = torch.stack([torch.stack([torch.randperm(nk, generator=g)[:n_sel] for _ in range(nq)]) for _ in range(heads)]).to(torch.int32)
    lens = torch.full((heads, nq), n_sel, dtype=torch.int32)
    exp = T._oracle_sparse(q, k, v, idx, lens, grid_q, grid_k, chunk, heads).float()
    qc, kc, vc, ic = q.to(cuda), k.to(cuda), v.to(cuda), idx.to(cuda)
    outs = []
    for rep in range(3):
        out = torch.full_like(q, float("nan")).to(cuda)
        lib.attention_bsa_bf16(qc, kc, vc, out, heads, ic, None, grid_q, grid_k, chunk)
        torch.cuda.synchronize()
        outs.append(out.cpu().float())
    print("case", case, "deterministic:", torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]))
    err = (outs[0] - exp).abs()
    print("  max err", err.max().item(), "nan", torch.isnan(outs[0]).sum().item())
    # per token chunk (in t,h,w order -> chunk id) and per head
    T_, H_, W_ = grid_q
    tok = torch.arange(T_ * H_ * W_)
    t_, h_, w_ = tok // (H_ * W_), (tok // W_) % H_, tok % W_
    cid = ((t_ // chunk[0]) * (H_ // chunk[1]) + h_ // chunk[1]) * (W_ // chunk[2]) + w_ // chunk[2]
    for hd in range(heads):
        e = err[:, hd * 128:(hd + 1) * 128].max(dim=1).values
        per = [round(e[cid == ci].max().item(), 3) for ci in range(nq)]
        print("  head", hd, "per-chunk max err", per[:16])
    bad = (err > 2e-2).nonzero()
    if len(bad):
        r = bad[0, 0].item()
        print("  first bad row", r, "chunk", cid[r].item(), "cols bad", (err[r] > 2e-2).sum().item(), "got", outs[0][r, :4].tolist(), "exp", exp[r, :4].tolist())
