#!/usr/bin/env python
"""Run the attack kernels and the 8x8-DCT masks kernel a few times on 1080p planes (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from b200wm import ops
import bench_extra as be
y = be.planes(64, 1080, 1920)
for _ in range(3):
    ops.attack_jpeg_requant_(y, 75)
    small = ops.attack_resize(y, (1280, 720), ops.INTER_AREA)
    ops.attack_resize(small, (1920, 1080), ops.INTER_LINEAR)
    ops.dct8_masks(y)
torch.cuda.synchronize()
print("ok")
