"""Device timings of the LongCat refine-pass attention at config-5 shape (16 x 44 x 80 latent tokens, 32 heads): gating
kernels, block-sparse attention, and the dense kernel on the same tensors (a development probe, not the bench)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from worldforge_b200 import lib

dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    T, H, W = int(os.environ.get("WF_T", 16)), int(os.environ.get("WF_H", 44)), int(os.environ.get("WF_W", 80))
    heads, C = 32, 4096
    res = {"grid": [T, H, W]}
    L = T * H * W
    qkv = torch.randn(L, 3 * C, device=dev).to(BF)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    out = torch.empty(L, C, device=dev, dtype=BF)
    dense = timeit(lambda: lib.attention_bf16(q, k, v, out, heads), iters=3, warm=1)
    res["dense"] = dict(ms=dense, tflops=4.0 * L * L * C / dense / 1e9)
    print("dense", res["dense"], flush=True)
    for chunk, sparsity in (((4, 4, 4), 0.9375), ((4, 4, 4), 0.875), ((4, 4, 8), 0.875), ((4, 4, 8), 0.75)):
        grid = (T, H, W)
        c = math.prod(chunk)
        if any(g % s for g, s in zip(grid, chunk)):
            continue
        nk = L // c
        n_sel = int((1 - sparsity) * nk)
        t_pool = timeit(lambda: (lib.bsa_mean_pool(q, grid, chunk, heads), lib.bsa_mean_pool(k, grid, chunk, heads)))
        qc, kc = lib.bsa_mean_pool(q, grid, chunk, heads), lib.bsa_mean_pool(k, grid, chunk, heads)
        t_sel = timeit(lambda: lib.bsa_select_topk(qc, kc, n_sel))
        idx = lib.bsa_select_topk(qc, kc, n_sel)
        t_att = timeit(lambda: lib.attention_bsa_bf16(q, k, v, out, heads, idx, None, grid, grid, chunk), iters=5, warm=2)
        useful = 4.0 * L * (n_sel * c) * C
        key = f"chunk{c}_sparsity{sparsity}"
        res[key] = dict(n_sel=n_sel, pool_ms=t_pool, select_ms=t_sel, attn_ms=t_att, useful_tflops=useful / t_att / 1e9,
                        kv_l2_gbps=heads * (L // c) * n_sel * c * 512 / t_att / 1e6, speedup_vs_dense=dense / (t_pool + t_sel + t_att))
        print(key, res[key], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bsa_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
