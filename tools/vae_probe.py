"""Device timings of the VAE round trip at the benchmark size (development probe, not the bench)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from worldforge_b200 import lib, vae as wvae

dev = torch.device("cuda:0")
F_, H, W = int(os.environ.get("WF_F", 81)), int(os.environ.get("WF_H", 480)), int(os.environ.get("WF_W", 832))
m = wvae.WfWanVAE.random_init(dev)
z = torch.randn(1, 16, (F_ - 1) // 4 + 1, H // 8, W // 8, device=dev)


def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        out = fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, out


# per-layer-kind accounting via events around the private layer methods
acc = {}
def wrap(name):
    orig = getattr(m, name)
    def f(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig(*a, **k); e.record()
        acc.setdefault(name, []).append((s, e, tuple(a[0].shape)))
        return r
    setattr(m, name, f)
for n in ("_res", "_attn", "_up", "_down"):
    wrap(n)

ms_d, vid = timed(lambda: m.decode(z)[0], 1)
acc.clear()
ms_d, vid = timed(lambda: m.decode(z)[0], 1)
torch.cuda.synchronize()
dec_break = {}
for k, v in acc.items():
    for s, e, shp in v:
        dec_break.setdefault(f"{k}{shp[1:]}", 0.0)
        dec_break[f"{k}{shp[1:]}"] += s.elapsed_time(e)
acc.clear()
ms_e, mu = timed(lambda: m.encode(vid).latent_dist.mode(), 1)
torch.cuda.synchronize()
enc_break = {}
for k, v in acc.items():
    for s, e, shp in v:
        enc_break.setdefault(f"{k}{shp[1:]}", 0.0)
        enc_break[f"{k}{shp[1:]}"] += s.elapsed_time(e)
res = dict(decode_ms=ms_d, encode_ms=ms_e, dec_tflops=275e12 * (F_ * H * W) / (81 * 480 * 832) / ms_d / 1e9,
           enc_tflops=164e12 * (F_ * H * W) / (81 * 480 * 832) / ms_e / 1e9, decode=dec_break, encode=enc_break,
           peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/vae_probe.json", "w"), indent=1)
