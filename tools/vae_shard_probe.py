"""Development probe: the row-sharded VAE at WF_P ranks evaluated rank by rank on ONE GPU (no communication): per stage the
slowest rank's device time, with and without the per-level cuts, against the unsharded time.  81 x 480 x 832 by default."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from worldforge_b200 import lib, vae as wvae

dev = torch.device("cuda:0")
lib.load()
P = int(os.environ.get("WF_P", 8))
F_, H, W = int(os.environ.get("WF_F", 81)), int(os.environ.get("WF_H", 480)), int(os.environ.get("WF_W", 832))
v = wvae.WfWanVAE.random_init(dev, seed=4321)
g = torch.Generator().manual_seed(0)
video = (torch.rand(1, 3, F_, H, W, generator=g) * 2 - 1).to(dev)
z = torch.randn(1, 16, (F_ - 1) // 4 + 1, H // 8, W // 8, generator=g).to(dev)


def t(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


res = {}
v.encode(video); v.decode(z)                                       # warm-up
res["unsharded_encode_ms"], mu = t(lambda: v.encode(video).latent_dist.mode())
res["unsharded_decode_ms"], dec = t(lambda: v.decode(z)[0])
for cuts in (False, True):
    v.level_cuts = cuts
    for which in ("enc", "dec"):
        if which == "enc":
            full = video[0]
        else:
            cl = lib.planar_to_cl(z[0].float().contiguous(), v.z_dim, round_tf32=True)
            full = v._conv(cl, "conv2", wvae.TAPS_1, v.z_dim, round_out=True)
        per_stage = []
        for stage in v.sharded_stages(which, full, P):
            parts, worst, tot = [], 0.0, 0.0
            for r in range(P):
                ms, (part, dim, bounds) = t(lambda: stage(r, full))
                parts.append(part); worst = max(worst, ms); tot += ms
            per_stage.append({"slowest_rank_ms": worst, "sum_ranks_ms": tot, "gather_mb": sum(p.numel() for p in parts) * 4 / 2 ** 20})
            full = wvae.assemble_rows(parts, dim)
        key = f"{which}_P{P}_cuts{int(cuts)}"
        res[key] = {"stages": per_stage, "critical_path_ms": sum(s["slowest_rank_ms"] for s in per_stage),
                    "work_ms": sum(s["sum_ranks_ms"] for s in per_stage)}
        if which == "enc":
            ok = torch.equal(lib.cl_to_planar(full, v.z_dim).unsqueeze(0), mu)
        else:
            ok = torch.equal(full.unsqueeze(0), dec)
        res[key]["bit_identical"] = bool(ok)
        print(key, "critical path", round(res[key]["critical_path_ms"], 1), "ms; total work", round(res[key]["work_ms"], 1),
              "ms; bit-identical", ok, "; gathers MB", [round(s["gather_mb"]) for s in per_stage], flush=True)
print(json.dumps({k: v_ for k, v_ in res.items() if not isinstance(v_, dict)}))
json.dump(res, open("gpurun_out/vae_shard_probe.json", "w"), indent=1)
