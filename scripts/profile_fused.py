#!/usr/bin/env python
"""Run the fused rgb24 kernels a few times on natural-image frames (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))
import numpy as np, torch
from b200wm import ops
from offmark_b200.generator.shuffler import Shuffler
DEV = torch.device("cuda:0")
h, w, n = 1080, 1920, 64
crop = np.load(os.path.join(ROOT, "tests", "golden", "frame63_crop.npz"))["bgr"]
tile = torch.from_numpy(np.tile(crop, (-(-h // crop.shape[0]), -(-w // crop.shape[1]), 1))[:h, :w].copy()).to(DEV)
g = torch.Generator(device=DEV).manual_seed(9)
nat = (tile[None].float() + 1.5 * torch.randn((n, h, w, 3), device=DEV, generator=g)).round().clamp(0, 255).to(torch.uint8)
wm, ln = ops.pack_bits(Shuffler(key=0).generate_wm(np.array([0, 1, 1, 0, 0, 1, 0, 1]), (1, h * w // 64))[0], device=DEV)
for _ in range(3):
    ops.dwtsvd_embed_rgb8_(nat, wm, ln)
    ops.dwtsvd_extract_rgb8(nat, payload_len=8)
torch.cuda.synchronize()
print("ok")
