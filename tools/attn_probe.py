"""Self-attention at the benchmark shape for the variant selected by WF_ATTN (development probe): device time, TFLOP/s,
and the largest difference against flash-attn 2 on the same inputs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from worldforge_b200 import lib

dev = torch.device("cuda:0")
L, H = int(os.environ.get("WF_L", 32760)), int(os.environ.get("WF_HEADS", 40))
D = H * 128
torch.manual_seed(0)
qkv = torch.randn(L, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(L, D, device=dev, dtype=torch.bfloat16)
fn = lambda: lib.attention_bf16(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, H)
fn(); torch.cuda.synchronize()
times = []
for _ in range(int(os.environ.get("WF_ITERS", 6))):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    times.append(s.elapsed_time(e))
ms = sorted(times)[len(times) // 2]
res = dict(variant=os.environ.get("WF_ATTN", "default"), L=L, heads=H, ms=ms, ms_min=min(times), tflops=4.0 * L * L * D / ms / 1e9)
try:
    from flash_attn import flash_attn_func
    q, k, v = (qkv[:, i * D:(i + 1) * D].reshape(1, L, H, 128) for i in range(3))
    o2 = flash_attn_func(q, k, v).reshape(L, D)
    res["maxdiff_vs_fa2"] = (o2.float() - out.float()).abs().max().item()
    res["rel_vs_fa2"] = ((o2.float() - out.float()).norm() / o2.float().norm()).item()
except Exception as ex:
    res["fa2"] = str(ex)
print(json.dumps(res))
