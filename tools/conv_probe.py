"""Development probe: one 3x3x3 causal convolution of the VAE at a full-resolution shape, by epilogue form (plain /
+ residual / + fused RMS-norm+SiLU with one or two outputs), CUDA-event timed.  WF_C (channels), WF_T, WF_H, WF_W."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from worldforge_b200 import lib

dev = torch.device("cuda:0")
lib.load()
Cc = int(os.environ.get("WF_C", 96)); T = int(os.environ.get("WF_T", 21)); H = int(os.environ.get("WF_H", 480)); W = int(os.environ.get("WF_W", 832))
g = torch.Generator(device=dev).manual_seed(1)
x = lib.round_tf32(torch.randn(T, H, W, Cc, device=dev, generator=g)) if hasattr(lib, "round_tf32") else torch.randn(T, H, W, Cc, device=dev, generator=g)
w = torch.randn(27 * Cc, Cc, device=dev, generator=g) * 0.02
b = torch.randn(Cc, device=dev, generator=g)
gamma = torch.randn(Cc, device=dev, generator=g)
res = torch.randn(T, H, W, Cc, device=dev, generator=g)
taps = [(dt - 2, dy - 1, dx - 1) for dt in range(3) for dy in range(3) for dx in range(3)]
out = torch.empty(T, H, W, Cc, device=dev)
out2 = torch.empty(T, H, W, Cc, device=dev)
flop = 2.0 * 27 * Cc * Cc * T * H * W


def timeit(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = {
    "plain": dict(),
    "resid": dict(resid=res),
    "norm_inplace": dict(norm_gamma=gamma),
    "norm_dual": dict(norm_gamma=gamma, norm_out=out2),
    "resid_norm_dual": dict(resid=res, norm_gamma=gamma, norm_out=out2),
}
resj = {"shape": [T, H, W, Cc]}
for name, kw in cases.items():
    ms = timeit(lambda: lib.conv_tf32(x, w, b, taps, out, T=T, H=H, W=W, Cout=Cc, tile_w=16, round_out=True, **kw))
    resj[name] = {"ms": ms, "tflops": flop / ms / 1e9}
    print(name, f"{ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s", flush=True)
    # per-role clocks of CTA 0 for one launch
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    lib._call("wf_debug_conv_profile", prof.data_ptr())
    lib.conv_tf32(x, w, b, taps, out, T=T, H=H, W=W, Cout=Cc, tile_w=16, round_out=True, **kw)
    torch.cuda.synchronize()
    lib._call("wf_debug_conv_profile", None)
    pr = prof.tolist()
    names = ["A-producer waits a_empty", "B-producer waits b_empty", "MMA waits acc_empty", "MMA waits a_full", "MMA waits b_full",
             "epilogue waits acc_full", "epilogue work", "total", "blocks"]
    nb = max(pr[8], 1)
    print("   per block (clocks): " + "; ".join(f"{n} {pr[i] / nb:.0f}" for i, n in enumerate(names[:8])) + f"; blocks {pr[8]}", flush=True)
    resj[name]["prof_per_block"] = {n: pr[i] / nb for i, n in enumerate(names[:8])}
ms = timeit(lambda: lib.rms_norm_cl(out, gamma, out=out2))
resj["rms_silu_separate_ms"] = ms
print("separate rms_silu pass", f"{ms:.3f} ms")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(resj, open("gpurun_out/conv_probe.json", "w"), indent=1)
