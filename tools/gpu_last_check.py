"""Round-2 last GPU call (budget: about a minute): the resize branch of fuse_latents on the device - a mis-sized clip / mask
gives exactly what the pre-sized pair gives."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wan_vae                                             # noqa: E402  (weights for the small VAE only)
from worldforge_b200 import scheduler as wsched, vae as wvae            # noqa: E402

dev = torch.device("cuda:0")
vae = wvae.WfWanVAE(wan_vae.init_params(wan_vae.VaeConfig(dim=8), 2), dev, dim=8)
g = torch.Generator().manual_seed(3)
x0 = torch.randn(1, 16, 2, 4, 6, generator=g).to(dev)
clip = torch.rand(1, 3, 5, 20, 30, generator=g).to(dev)
mask = (torch.rand(1, 3, 5, 20, 30, generator=g) > 0.4).float().to(dev)
s = wsched.WfUniPCScheduler(flow_shift=3.0)
s.set_timesteps(4, device=dev)
a = s.fuse_latents(x0, clip, mask, vae=vae)
b = s.fuse_latents(x0, clip, mask, vae=vae)                             # second call: the kept pair
clip_s = F.interpolate(clip.reshape(15, 1, 20, 30), size=(32, 48), mode="bilinear", align_corners=False).reshape(1, 3, 5, 32, 48)
mask_s = F.interpolate(mask[:, 0:1].reshape(5, 1, 20, 30), size=(32, 48), mode="nearest").reshape(1, 1, 5, 32, 48)
c = wsched.WfUniPCScheduler(flow_shift=3.0).fuse_latents(x0, clip_s.contiguous(), mask_s.contiguous(), vae=vae)
torch.cuda.synchronize()
assert torch.equal(a, b) and torch.equal(a, c) and not torch.equal(a, x0), "resize branch differs from the pre-sized pair"
print("presize on the device: ok", tuple(a.shape))
