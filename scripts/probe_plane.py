#!/usr/bin/env python
"""Embed / extract kernel times of uint8 planes on one GPU:  python scripts/probe_plane.py [H W [frames]]  (default 4K).
A/B aid for the strip shapes of the TMA kernels (B200WM_NO_EMBED_WIDE=1: 4K embed back to column chunks).  Prints ms, GB/s, fraction of the measured HBM
peak and a checksum of the marked planes and votes (must not depend on the strip shape)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))
import numpy as np, torch
from b200wm import ops

I420 = "i420" in sys.argv                    # Y planes inside planar 4:2:0 frames (frame stride W*H*3/2), as bench.py holds them
args = [int(a) for a in sys.argv[1:] if a.isdigit()]
H, W = (args + [2160, 3840])[:2] if len(args) >= 2 else (2160, 3840)
N = args[2] if len(args) > 2 else max(64, int(256 * 2160 * 3840 / (H * W)))
dev = torch.device("cuda:0")
peak = 6552.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
g = torch.Generator(device=dev).manual_seed(7)
xx = torch.arange(W, device=dev, dtype=torch.float32)[None, None, :]
yy = torch.arange(H, device=dev, dtype=torch.float32)[None, :, None]
if I420:
    src = ops.i420_plane(torch.full((N, H * W * 3 // 2), 128, dtype=torch.uint8, device=dev), H, W, "y")
    dst = ops.i420_plane(torch.full((N, H * W * 3 // 2), 128, dtype=torch.uint8, device=dev), H, W, "y")
else:
    src = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
for f0 in range(0, N, 16):
    f = torch.arange(f0, min(N, f0 + 16), device=dev, dtype=torch.float32)[:, None, None]
    y = 128 + 80 * torch.sin(2 * np.pi * (3 * xx / W + f / 97)) * torch.cos(2 * np.pi * (2 * yy / H + f / 53))
    src[f0:f0 + 16] = (y + 6 * torch.randn(y.shape, device=dev, generator=g)).round().clamp(16, 235).to(torch.uint8)
rng = np.random.RandomState(1)
ROWS = "rows" in sys.argv                    # one watermark row per 60-frame segment, picked per frame (bench.py's batch)
wm, ln = ops.pack_bits(rng.randint(0, 2, ((N + 59) // 60 if ROWS else 1, H * W // 64)), device=dev)
frame_row = (torch.arange(N, device=dev, dtype=torch.int32) // 60) if ROWS else None


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    z.record()
    torch.cuda.synchronize()
    return a.elapsed_time(z) / reps


e = timed(lambda: ops.dwtsvd_embed_(src, wm, ln, frame_wm_row=frame_row, out=dst))
x = timed(lambda: ops.dwtsvd_extract(dst, payload_len=8))
# the two kernels alternating, as a step of bench.py runs them (per-kernel events)
marks = []
for rep in range(13):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    ops.dwtsvd_embed_(src, wm, ln, frame_wm_row=frame_row, out=dst)
    ev[1].record()
    ops.dwtsvd_extract(dst, payload_len=8)
    ev[2].record()
    if rep >= 3:
        marks.append(ev)
torch.cuda.synchronize()
e_alt = float(np.mean([m[0].elapsed_time(m[1]) for m in marks]))
x_alt = float(np.mean([m[1].elapsed_time(m[2]) for m in marks]))
raw, counts = ops.dwtsvd_extract(dst, payload_len=8)
weights = torch.arange(1, 1 + H * W, device=dev, dtype=torch.int64).view(H, W) % 65521
check = int((dst[:8].to(torch.int64) * weights).sum()) ^ int(counts.to(torch.int64).sum())
gb = N * H * W / 1e9
print(json.dumps({"plane": [H, W], "frames": N, "i420": I420, "rows": ROWS, "env": {k: v for k, v in os.environ.items() if k.startswith("B200WM_")}, "embed_ms": round(e, 4), "embed_GBs": round(2 * gb / e * 1e3, 1),
                  "embed_frac": round(2 * gb / e * 1e3 / peak, 4), "extract_ms": round(x, 4), "extract_frac": round(gb / x * 1e3 / peak, 4),
                  "alternating": {"embed_ms": round(e_alt, 4), "extract_ms": round(x_alt, 4)}, "checksum": check}))
