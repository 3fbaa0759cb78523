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


if __name__ == "__main__":
    main()
