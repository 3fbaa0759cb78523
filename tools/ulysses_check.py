"""torchrun --nproc-per-node 2 tools/ulysses_check.py : sequence-parallel DiT forward == single-GPU forward (tiny widths)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from worldforge_b200 import transformer as wtr, ulysses

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
cfg = wtr.WanDitConfig(dim=512, ffn_dim=1024, num_heads=4, num_layers=2, text_dim=64, text_len=16, img_dim=64, img_len=5, freq_dim=32)
m = wtr.WfWanTransformer.random_init(cfg, dev, seed=5)
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 36, 3, 8, 16, generator=g).to(torch.bfloat16).to(dev)
ctx = torch.randn(1, 16, 64, generator=g).to(torch.bfloat16).to(dev)
clip = torch.randn(1, 5, 64, generator=g).to(torch.bfloat16).to(dev)
t = torch.tensor([500], device=dev)
single = m(x, t, ctx, clip)[0].clone()
sp = ulysses.enable(m)
par = m(x, t, ctx, clip)[0]
torch.cuda.synchronize()
same = torch.equal(single, par)
md = (single.float() - par.float()).abs().max().item()
print(f"rank {rank}/{world}: sequence-parallel forward equal to single-GPU: {same} (max diff {md}), a2a calls {sp.a2a_calls}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if md < 1e-2 else 1)
