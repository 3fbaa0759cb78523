"""Quick device timings of the tensor-core kernels at Wan-14B 480p shapes (not the bench; a development probe)."""
import sys, os, json, time
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
    L = int(os.environ.get("WF_L", 32760))
    D, Fd, H = 5120, 13824, 40
    res = {}
    a = torch.randn(L, D, device=dev).to(BF)
    for name, N, K, epi in [("qkv", 3 * D, D, 0), ("o_resid", D, D, 2), ("ffn0_gelu", Fd, D, 1), ("ffn2_resid", D, Fd, 2)]:
        A = torch.randn(L, K, device=dev).to(BF) if K != D else a
        W = (torch.randn(N, K, device=dev) * 0.02).to(BF)
        b = torch.zeros(N, device=dev, dtype=BF)
        out = torch.zeros(L, N, device=dev, dtype=(torch.float32 if epi == 2 else BF))
        gate = torch.ones(N, device=dev)
        ms = timeit(lambda: lib.gemm_bf16(A, W, b, out, epi, gate=gate if epi == 2 else None))
        fl = 2.0 * L * N * K
        ref = timeit(lambda: torch.matmul(A, W.t()))
        res[name] = dict(ms=ms, tflops=fl / ms / 1e9, cublas_ms=ref, cublas_tflops=fl / ref / 1e9)
        print(name, res[name], flush=True)
        del A, W, out
    qkv = torch.randn(L, 3 * D, device=dev).to(BF)
    out = torch.empty(L, D, device=dev, dtype=BF)
    ms = timeit(lambda: lib.attention_bf16(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, H), iters=3, warm=1)
    fl = 4.0 * L * L * D
    res["self_attn"] = dict(ms=ms, tflops=fl / ms / 1e9)
    print("self_attn", res["self_attn"], flush=True)
    try:
        from flash_attn import flash_attn_func
        q = qkv[:, :D].reshape(1, L, H, 128); k = qkv[:, D:2 * D].reshape(1, L, H, 128); v = qkv[:, 2 * D:].reshape(1, L, H, 128)
        ms2 = timeit(lambda: flash_attn_func(q, k, v), iters=3, warm=1)
        res["flash_attn2"] = dict(ms=ms2, tflops=fl / ms2 / 1e9)
        o2 = flash_attn_func(q, k, v).reshape(L, D)
        res["attn_vs_fa2_maxdiff"] = (o2.float() - out.float()).abs().max().item()
        print("flash_attn2", res["flash_attn2"], "maxdiff", res["attn_vs_fa2_maxdiff"], flush=True)
    except Exception as ex:  # noqa
        print("flash_attn unavailable:", ex)
    kv = torch.randn(769, 2 * D, device=dev).to(BF)
    ms = timeit(lambda: lib.attention_bf16(qkv[:, :D], kv[:512, :D], kv[:512, D:], out, H))
    res["cross_attn_512"] = dict(ms=ms, tflops=4.0 * L * 512 * D / ms / 1e9)
    print("cross512", res["cross_attn_512"], flush=True)
    x = torch.randn(L, D, device=dev); h = torch.empty(L, D, device=dev, dtype=BF)
    sc = torch.zeros(D, device=dev); 
    ms = timeit(lambda: lib.layer_norm(x, h, 1e-6, scale=sc, shift=sc))
    res["ln_mod"] = dict(ms=ms, gbs=L * D * 6 / ms / 1e6)
    w = torch.ones(D, device=dev)
    ms = timeit(lambda: lib.rms_norm_rope_(qkv[:, :D], w, 1e-6, None))
    res["rms"] = dict(ms=ms, gbs=L * D * 4 / ms / 1e6)
    print(res["ln_mod"], res["rms"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
