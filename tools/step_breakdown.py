"""Where one guided step of the benchmark configuration spends its time (development probe, not the bench).

Times the pieces of a guided step at Wan2.1-I2V-14B 480p / 81 frames separately, each bracketed by a device
synchronise: one DiT forward, one VAE decode, one VAE encode, the FLF blend, the FLF channel scoring (GPU
quantisation + D2H + host Farneback + metrics; wall clock), and the CPU-generator noise draw + upload of IRR.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from worldforge_b200 import flf_select, lib, synth, transformer as wtr, vae as wvae

dev = torch.device("cuda:0")
layers = int(os.environ.get("WF_LAYERS", 40))
F_, H, W = 81, 480, 832
lib.load()
tr = wtr.WfWanTransformer.random_init(wtr.WanDitConfig(num_layers=layers), dev, seed=1234)
vae = wvae.WfWanVAE.random_init(dev, seed=4321)
inp = synth.make_inputs(F_, H, W, seed=42)
d = {k: getattr(inp, k).to(dev) for k in ("latents", "condition", "prompt_embeds", "negative_prompt_embeds", "image_embeds",
                                           "video_ref", "mask")}


def wall(fn, n=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


res = {}
model_in = torch.cat([d["latents"], d["condition"]], dim=1).to(torch.bfloat16)
t = torch.tensor([900], device=dev)
res["dit_forward_ms"], v = wall(lambda: tr(hidden_states=model_in, timestep=t, encoder_hidden_states=d["prompt_embeds"],
                                           encoder_hidden_states_image=d["image_embeds"], return_dict=False)[0])
z = d["latents"].clone()
res["vae_decode_ms"], dec = wall(lambda: vae.decode(z, return_dict=False)[0])
res["flf_blend_ms"], fused = wall(lambda: lib.flf_blend(dec.contiguous(), d["video_ref"], d["mask"]))
res["vae_encode_ms"], enc = wall(lambda: vae.encode(fused).latent_dist.mode())
sel = flf_select.FlowChannelSelector()
x0 = d["latents"].to(torch.bfloat16)
res["flf_scores_ms"], sc = wall(lambda: sel.scores(x0, enc.to(torch.bfloat16)))
res["flf_threads"] = sel.threads
g = torch.Generator().manual_seed(1)
res["irr_noise_ms"], _ = wall(lambda: torch.randn(x0.shape, generator=g).pin_memory().to(dev, non_blocking=True))
res["host_cores"] = os.cpu_count()
res["guided_step_estimate_ms"] = 4 * res["dit_forward_ms"] + 2 * (res["vae_decode_ms"] + res["vae_encode_ms"] + res["flf_blend_ms"]) \
    + res["flf_scores_ms"] + res["irr_noise_ms"]
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/step_breakdown.json", "w"), indent=1)
