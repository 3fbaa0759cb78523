"""Writes tests/golden/inputs_golden.pt: outputs of the reference's own ``soften_mask`` (infer_worldforge.py:105-150, taken
from the script with ``ast``) on seeded masks.  Run where /root/reference is mounted:  python -m oracle.make_inputs_golden"""
import os

import numpy as np
import torch

from oracle import inputs as oin

REF = "/root/reference/wan_for_worldforge/infer_worldforge.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs_golden.pt")


def masks(seed: int = 0) -> np.ndarray:
    """uint8 [6, 70, 90]: blobs, a half plane, thin structures, an all-set and an all-unset frame, fractional edge values."""
    rng = np.random.default_rng(seed)
    F, H, W = 6, 70, 90
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((F, H, W), np.uint8)
    m[0] = (((yy - 30) ** 2 + (xx - 40) ** 2) < 20 ** 2) * 255
    m[1] = (xx > 37) * 255
    m[1, 10:12, :] = 0; m[1, :, 60] = 0
    m[2] = (rng.random((H, W)) > 0.02) * 255
    m[3] = 255
    m[4] = 0
    m[5] = (yy + xx > 60) * 255
    m[5][(yy + xx > 60) & (yy + xx < 64)] = 128                    # resized masks carry in-between values
    return m


def main():
    ref = oin.reference_soften_mask(REF)
    mu8 = masks()
    arr = oin.stack_masks(mu8)
    gold = {"masks_u8": torch.from_numpy(mu8)}
    for td, kind in ((15, "sine"), (7, "linear"), (10, "exponential"), (4, "cosine")):
        gold[f"soft_{td}_{kind}"] = torch.from_numpy(ref(arr, td, kind))
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__" and len(__import__("sys").argv) == 1:
    main()


# ---------------------------------------------------------------------------------------------------------------------
# LongCat video continuation (KV cache): python -m oracle.make_inputs_golden kv  ->  tests/golden/longcat_kv_golden.pt
# ---------------------------------------------------------------------------------------------------------------------
KV_OUT = os.path.join(os.path.dirname(OUT), "longcat_kv_golden.pt")


def kv_inputs():
    g = torch.Generator().manual_seed(4)
    lat = torch.randn(1, 16, 5, 8, 12, generator=g)               # 2 clean condition frames + 3 noise frames
    ctx = torch.randn(1, 1, 10, 64, generator=g)
    mask = torch.ones(1, 10, dtype=torch.int64); mask[:, 7:] = 0
    return lat, ctx, mask


def ref_kv():
    """The reference DiT's caching pass (return_kv, skip_crs_attn; pipeline_longcat_video.py:336-350) on the condition
    frames and its forward over the noise frames with that cache (:1202-1209), fp32."""
    from oracle import longcat_dit, make_golden as mg, ref_shim
    cfg = longcat_dit.LongCatConfig(**mg.LC_DIT)
    P = longcat_dit.init_params(cfg, 3)
    mod = ref_shim.load_longcat_dit_module()
    m = mod.LongCatVideoTransformer3DModel(in_channels=16, out_channels=16, hidden_size=256, depth=2, num_heads=2, caption_channels=64,
                                           mlp_ratio=4, adaln_tembed_dim=32, frequency_embedding_size=32, enable_xformers=True,
                                           cp_split_hw=[1, 1]).eval()
    m.load_state_dict({k: P[k].clone() for k in m.state_dict()})
    lat, ctx, mask = kv_inputs()
    with torch.no_grad():
        out_c, cache = m(lat[:, :, :2], torch.zeros(1, 2), torch.zeros(1, 1, 10, 64), return_kv=True, skip_crs_attn=True)
        out_n = m(lat[:, :, 2:], torch.full((1, 3), 600.0), ctx, encoder_attention_mask=mask, num_cond_latents=2, kv_cache_dict=cache)
        joint = m(lat, torch.tensor([[0.0, 0.0, 600.0, 600.0, 600.0]]), ctx, encoder_attention_mask=mask, num_cond_latents=2)
    return dict(cond_out=out_c[0].clone(), noise_out=out_n[0].clone(), joint_noise_out=joint[0][:, 2:].clone(),
                k0=cache[0][0][0].clone(), v1=cache[1][1][0].clone())


if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "kv":
    torch.save(ref_kv(), KV_OUT)
    print("wrote", KV_OUT, os.path.getsize(KV_OUT), "bytes")
