#!/usr/bin/env python
"""BASELINE config 5: detection after distortions over 10,000 marked 1080p frames.

All frames are generated, marked, attacked and read back on the GPU (b200wm kernels); on a
subsample per attack the reference extractor (oracle restatement, CPU) reads the very same attacked
frames so that the two bit-error rates can be put side by side.  Prints a markdown table.
    python tests/studies/attack_study.py [--frames 10000] [--oracle-frames 64]

Lives under tests/ because it runs the oracle (as the checker) next to the CUDA path: only tests/, smoke() and
bench.py's CPU legs may do that.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))

import numpy as np      # noqa: E402
import torch            # noqa: E402

from b200wm import ops  # noqa: E402
from offmark_b200.generator.shuffler import Shuffler            # noqa: E402
from offmark_b200.degenerator.de_shuffler import DeShuffler     # noqa: E402
from oracle import dwt_dct_svd as o_svd, payload as o_pay       # noqa: E402  (checker only)

H, W, KEY = 1080, 1920, 0
PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])


def make_frames(n, dev):
    out = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(2026)
    xx = torch.arange(W, device=dev, dtype=torch.float32)[None, None, :]
    yy = torch.arange(H, device=dev, dtype=torch.float32)[None, :, None]
    for f0 in range(0, n, 50):
        m = min(50, n - f0)
        f = torch.arange(f0, f0 + m, device=dev, dtype=torch.float32)[:, None, None]
        y = 128 + 80 * torch.sin(2 * np.pi * (3 * xx / W + f / 97)) * torch.cos(2 * np.pi * (2 * yy / H + f / 53))
        out[f0:f0 + m] = (y + 6 * torch.randn((m, H, W), device=dev, generator=g)).round().clamp(16, 235).to(torch.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--oracle-frames", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, block_num = args.frames, H * W // 64
    wm = Shuffler(key=KEY).generate_wm(PAYLOAD, (1, block_num))
    packed, ln = ops.pack_bits(wm[0], device=dev)
    truth = torch.from_numpy(wm[0].astype(np.uint8)).to(dev)
    deg = DeShuffler(key=KEY).set_shape((8,))

    t0 = time.time()
    marked = make_frames(n, dev)
    ops.dwtsvd_embed_(marked, packed, ln)
    torch.cuda.synchronize()
    print(f"generated and marked {n} frames in {time.time() - t0:.1f} s", file=sys.stderr)
    work = torch.empty_like(marked)

    def noise(sigma):
        g = torch.Generator(device=dev).manual_seed(int(sigma * 1000))
        for f0 in range(0, n, 100):
            m = min(100, n - f0)
            ops.attack_add_noise_(work[f0:f0 + m], sigma * torch.randn((m, H, W), device=dev, generator=g))

    def resize():
        for f0 in range(0, n, 500):
            ops.attack_resize_roundtrip_(work[f0:f0 + 500])

    attacks = [("none", lambda: None), ("jpeg-like q95", lambda: ops.attack_jpeg_requant_(work, 95)),
               ("jpeg-like q85", lambda: ops.attack_jpeg_requant_(work, 85)),
               ("jpeg-like q75", lambda: ops.attack_jpeg_requant_(work, 75)),
               ("gaussian sigma 1", lambda: noise(1.0)), ("gaussian sigma 2", lambda: noise(2.0)),
               ("gaussian sigma 4", lambda: noise(4.0)),
               ("resize 1080p-720p-1080p (area, bilinear)", lambda: resize())]
    rows = []
    for name, attack in attacks:
        work.copy_(marked)
        attack()
        raw, counts = ops.dwtsvd_extract(work, payload_len=8)
        patterns, _ = deg.degenerate_counts(counts, block_num)
        torch.cuda.synchronize()
        # raw BER vs the embedded bits, on the GPU (plumbing: unpack with torch ops)
        bits = ((raw.view(torch.uint8)[:, :, None] >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).reshape(n, -1)[:, :block_num]
        ber_gpu = float((bits != truth[None, :]).float().mean())
        payload_ok = float((patterns == torch.from_numpy(PAYLOAD.astype(np.uint8)).to(dev)).all(dim=1).float().mean())
        # reference extractor on a subsample of the same attacked frames
        k = min(args.oracle_frames, n)
        idx = np.linspace(0, n - 1, k).astype(int)
        host = work[torch.from_numpy(idx).to(dev)].cpu().numpy()
        ber_ref, agree, pay_same = [], [], []
        for j, f in enumerate(idx):
            ref_bits = o_svd.extract_plane(host[j])[0].astype(np.uint8)
            gpu_bits = bits[f].cpu().numpy()
            ber_ref.append((ref_bits != wm[0]).mean())
            agree.append((ref_bits == gpu_bits).mean())
            pay_same.append(np.array_equal(o_pay.degenerate(ref_bits.reshape(1, -1), 8, KEY), patterns[f].cpu().numpy()))
        rows.append({"attack": name, "frames": n, "raw_ber_b200": ber_gpu, "payload_exact_b200": payload_ok,
                     "oracle_frames": k, "raw_ber_reference": float(np.mean(ber_ref)),
                     "raw_bit_agreement": float(np.mean(agree)), "payload_identical_to_reference": float(np.mean(pay_same))})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print("| attack | frames | raw BER (B200 extractor, all frames) | payload exact (all frames) | raw BER (reference extractor, subsample) | raw-bit agreement B200 vs reference | voted payload identical |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['attack']} | {r['frames']} | {r['raw_ber_b200']:.5f} | {r['payload_exact_b200']:.4f} | "
              f"{r['raw_ber_reference']:.5f} ({r['oracle_frames']} frames) | {r['raw_bit_agreement']:.6f} | {r['payload_identical_to_reference']:.3f} |")


if __name__ == "__main__":
    main()
