#!/usr/bin/env python
"""Time the three 8x8-DCT kernels on planar uint8 1080p planes (fractions of the measured HBM peak)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
from b200wm import ops
from offmark_b200.generator.shuffler import Shuffler
import bench_extra as be
h, w, n = 1080, 1920, 512
yp, up = be.planes(n, h, w, seed=5), be.planes(n, h, w, seed=6)
wm, ln = ops.pack_bits(Shuffler(key=0).generate_wm(np.array([0, 1, 1, 0, 0, 1, 0, 1]), (1, h * w // 64))[0], device=be.DEV)
ms_m = be.timed(lambda: ops.dct8_masks(yp)); masks = ops.dct8_masks(yp)
ms_e = be.timed(lambda: ops.dct8_embed_(up, masks, wm, ln, alpha=20))
ms_x = be.timed(lambda: ops.dct8_extract(up, masks, alpha=20, payload_len=8))
f = lambda ms, b: round(n * b * h * w / (ms * 1e-3) / 1e9 / be.PEAK, 3)
print("masks", f(ms_m, 1), "embed", f(ms_e, 2), "extract", f(ms_x, 1))
