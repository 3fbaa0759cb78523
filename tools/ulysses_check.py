"""torchrun --nproc-per-node N tools/ulysses_check.py : sequence-parallel DiT forward == single-GPU forward (tiny widths),
row-sharded VAE encode / decode == single-GPU encode / decode (bit for bit), rank-sharded FLF scoring == single-rank scoring."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from worldforge_b200 import transformer as wtr, ulysses

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
nh = max(4, world)                                # heads and tokens must split over the ranks
cfg = wtr.WanDitConfig(dim=128 * nh, ffn_dim=1024, num_heads=nh, num_layers=2, text_dim=64, text_len=16, img_dim=64, img_len=5, freq_dim=32)
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
sp.peer = True                                   # the same exchange through NVLink peer memory (no collective on the data path)
par2 = m(x, t, ctx, clip)[0]
par3 = m(x, t, ctx, clip)[0]                     # second forward: buffers and layer parity are reused
torch.cuda.synchronize()
psp = next(iter(m._psp.values()))
peer_ok = torch.equal(single, par2) and torch.equal(single, par3)
print(f"rank {rank}/{world}: peer-memory forward equal to single-GPU: {peer_ok} (max diff {(single.float() - par2.float()).abs().max().item()}), "
      f"barriers {psp.barriers}, a2a calls still {sp.a2a_calls}", flush=True)
md = max(md, 0.0 if peer_ok else 1.0)
# CFG x Ulysses layout: the two halves of the ranks run the conditional / unconditional forward (Ulysses inside a half, over
# peer memory) and swap the predictions - both must equal the single-GPU forwards bit for bit
ctx_u = torch.randn(1, 16, 64, generator=g).to(torch.bfloat16).to(dev)
m.sp = None
single_u = m(x, t, ctx_u, clip)[0].clone()
cfg_ok = True
if world % 2 == 0:
    sp_group, cfgp = ulysses.cfg_layout(world, rank)
    m._psp = {}
    if sp_group is not None:
        ulysses.enable(m, sp_group, peer=True)
    mine = m(x, t, ctx_u if cfgp.branch else ctx, clip)[0]
    v_c, v_u = cfgp.exchange(mine)
    torch.cuda.synchronize()
    cfg_ok = torch.equal(v_c, single) and torch.equal(v_u, single_u)
print(f"rank {rank}/{world}: cfg-parallel forwards equal to single-GPU: {cfg_ok}", flush=True)
md = max(md, 0.0 if cfg_ok else 1.0)
from worldforge_b200 import flf_select, vae as wvae
v = wvae.WfWanVAE.random_init(dev, dim=16, seed=3)
video = (torch.rand(1, 3, 9, 128, 96, generator=g) * 2 - 1).to(dev)
z = torch.randn(1, 16, 3, 16, 12, generator=g).to(dev)
mu1, dec1 = v.encode(video).latent_dist.mode().clone(), v.decode(z)[0].clone()
v.enable_row_sharding(dist.group.WORLD)
mu2, dec2 = v.encode(video).latent_dist.mode(), v.decode(z)[0]
torch.cuda.synchronize()
vae_ok = torch.equal(mu1, mu2) and torch.equal(dec1, dec2)
print(f"rank {rank}/{world}: row-sharded VAE equal to single-GPU: encode {torch.equal(mu1, mu2)} decode {torch.equal(dec1, dec2)}", flush=True)
a, b = torch.randn(1, 16, 6, 30, 52, generator=g).to(dev), torch.randn(1, 16, 6, 30, 52, generator=g).to(dev)
s1 = flf_select.FlowChannelSelector().scores(a, b)
s2 = flf_select.FlowChannelSelector(group=dist.group.WORLD, world=world, rank=rank).scores(a, b)
print(f"rank {rank}/{world}: rank-sharded FLF scores equal: {s1 == s2}", flush=True)
# LongCat context parallel with the reference's 2-D split (context_parallel_util.py:91-121, 231-243): every rank keeps one
# block of every frame, heads <-> tokens around self-attention, output blocks gathered - against the single-GPU forward (dense
# attention: same math with the keys in rank-major order, so equal up to bf16 rounding flips, not bit for bit)
from worldforge_b200 import longcat
split = min(([i, world // i] for i in range(1, int(world ** 0.5) + 1) if world % i == 0), key=lambda f: abs(f[0] - f[1]))
lheads = max(2, world)
lcfg = longcat.LongCatConfig(hidden_size=128 * lheads, depth=2, num_heads=lheads, caption_channels=32, adaln_tembed_dim=32, frequency_embedding_size=32)
lm = longcat.WfLongCatTransformer.random_init(lcfg, dev, seed=7)
lx = torch.randn(1, 16, 3, 16, 32, generator=g).to(dev)
lts = torch.tensor([[0.0, 700.0, 700.0]], device=dev)
lctx = torch.randn(1, 1, 8, 32, generator=g).to(torch.bfloat16).to(dev)
lmask = torch.ones(1, 8, dtype=torch.int64, device=dev); lmask[0, 6:] = 0
l_single = lm(lx, lts, lctx, encoder_attention_mask=lmask, num_cond_latents=1).clone()
lm.enable_context_parallel(dist.group.WORLD, split)
l_cp = lm(lx, lts, lctx, encoder_attention_mask=lmask, num_cond_latents=1)
torch.cuda.synchronize()
l_rel = ((l_cp - l_single).norm() / l_single.norm()).item()
same_everywhere = [torch.empty_like(l_cp) for _ in range(world)]
dist.all_gather(same_everywhere, l_cp)
l_ok = l_rel < 5e-3 and all(torch.equal(t_, l_cp) for t_ in same_everywhere)
print(f"rank {rank}/{world}: LongCat context-parallel forward (split {split[0]}x{split[1]}) matches single-GPU: {l_ok} (rel {l_rel:.2e})", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if md < 1e-2 and vae_ok and s1 == s2 and l_ok else 1)
