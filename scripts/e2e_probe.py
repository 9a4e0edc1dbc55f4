#!/usr/bin/env python
"""Where the end-to-end (host-buffer) path spends its time: per call and per chunk size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))
import numpy as np, torch
from b200wm import ops
from offmark_b200.generator.shuffler import Shuffler
from offmark_b200.degenerator.de_shuffler import DeShuffler
H, W, N = 1080, 1920, int(os.environ.get("N", "3000"))
dev = torch.device("cuda:0")
host_in = torch.empty((N, H, W), dtype=torch.uint8, pin_memory=True)
host_out = torch.empty((N, H, W), dtype=torch.uint8, pin_memory=True)
g = torch.Generator(device=dev).manual_seed(1)
for f0 in range(0, N, 250):
    host_in[f0:f0 + 250].copy_(torch.randint(16, 236, (min(250, N - f0), H, W), dtype=torch.uint8, device=dev, generator=g))
rows = np.stack([Shuffler(key=0).generate_wm(np.array([int(b) for b in format(s, "08b")]), (1, H * W // 64))[0] for s in range(N // 60 + 1)])
frame_row = (np.arange(N) // 60).astype(np.int32)
perm = DeShuffler(key=0).set_shape((8,)).payload_idx
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t0 = time.perf_counter(); packed, n = ops.pack_bits(rows); print("pack_bits of the payload rows: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
gb = N * H * W / 1e9
for chunk in (0, 16, 32, 64, 125, 250):
    tm = t(lambda: ops.dwtsvd_mark_host(host_in, host_out, rows, frame_wm_row=frame_row, chunk_frames=chunk))
    td = t(lambda: ops.dwtsvd_detect_host(host_out, perm, chunk_frames=chunk))
    print(f"chunk {chunk:4d}: mark {tm*1e3:7.1f} ms ({gb/tm:5.1f} GB/s each way)  detect {td*1e3:7.1f} ms ({gb/td:5.1f} GB/s)  -> {N/(tm+td):8.0f} frames/s")
