#!/usr/bin/env python
"""Secondary measurements (not the headline bench line): 4K planes, the fused rgb24 path, the 8x8
DCT pair and the attack kernels, each timed with CUDA events on HBM-resident data larger than L2.
Prints one JSON object per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))

import numpy as np      # noqa: E402
import torch            # noqa: E402

from b200wm import ops  # noqa: E402
from offmark_b200.generator.shuffler import Shuffler  # noqa: E402

DEV = torch.device("cuda:0")
PEAK = 6552.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK)


def timed(fn, warmup=3, steps=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def report(name, ms, frames, bytes_per_frame, **extra):
    gbs = frames * bytes_per_frame / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "ms": round(ms, 4), "frames": frames, "frames_per_s": round(frames / (ms * 1e-3)),
                      "algorithmic_GBs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 3), **extra}), flush=True)


def planes(n, h, w, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    base = 128 + 60 * torch.sin(torch.arange(w, device=DEV) / 41.0)[None, None, :] * torch.cos(torch.arange(h, device=DEV) / 29.0)[None, :, None]
    out = torch.empty((n, h, w), dtype=torch.uint8, device=DEV)
    for f0 in range(0, n, 32):
        m = min(32, n - f0)
        out[f0:f0 + m] = (base + 6 * torch.randn((m, h, w), device=DEV, generator=g)).round().clamp(16, 235).to(torch.uint8)
    return out


def main():
    payload = np.array([0, 1, 1, 0, 0, 1, 0, 1])
    # ---- 4K Y planes (BASELINE config 3, one GPU's shard)
    h, w, n = 2160, 3840, 256
    src = planes(n, h, w)
    dst = src.clone()
    wm, ln = ops.pack_bits(Shuffler(key=0).generate_wm(payload, (1, h * w // 64))[0], device=DEV)
    for path, tag in ((0, "tma"), (1, "ldg")):
        ops.set_path(path)
        ms_e = timed(lambda: ops.dwtsvd_embed_(src, wm, ln, out=dst))
        ms_x = timed(lambda: ops.dwtsvd_extract(dst, payload_len=8))
        report(f"4K embed ({tag})", ms_e, n, 2 * h * w)
        report(f"4K extract ({tag})", ms_x, n, h * w)
        report(f"4K embed+extract ({tag})", ms_e + ms_x, n, 3 * h * w)
    ops.set_path(0)
    del src, dst
    # ---- portrait 1080 x 1920 luma (135 tiles per row, rows only 8-byte aligned: whole-strip bulk copies still apply)
    hp, wp, n = 1920, 1080, 1024
    src = planes(n, hp, wp)
    dst = src.clone()
    wmp, lnp = ops.pack_bits(Shuffler(key=0).generate_wm(payload, (1, hp * wp // 64))[0], device=DEV)
    for path, tag in ((0, "tma"), (1, "ldg")):
        ops.set_path(path)
        ms_e = timed(lambda: ops.dwtsvd_embed_(src, wmp, lnp, out=dst))
        ms_x = timed(lambda: ops.dwtsvd_extract(dst, payload_len=8))
        report(f"portrait 1080x1920 embed+extract ({tag})", ms_e + ms_x, n, 3 * hp * wp)
    ops.set_path(0)
    del src, dst
    # ---- chroma plane of yuv420p 1080p frames, marked in place through a strided view
    h, w, n = 1080, 1920, 1024
    frames = torch.empty((n, h * w * 3 // 2), dtype=torch.uint8, device=DEV)
    frames[:, :h * w] = planes(n, h, w).view(n, -1)
    frames[:, h * w:] = planes(n, h // 2, w, seed=3).view(n, -1)
    u = ops.i420_plane(frames, h, w, "u")
    wmu, lnu = ops.pack_bits(Shuffler(key=0).generate_wm(payload, (1, (h // 2) * (w // 2) // 64))[0], device=DEV)
    for path, tag in ((0, "tma"), (1, "ldg")):
        ops.set_path(path)
        ms_e = timed(lambda: ops.dwtsvd_embed_(u, wmu, lnu))
        ms_x = timed(lambda: ops.dwtsvd_extract(u, payload_len=8))
        report(f"I420 U plane 960x540 embed in place ({tag})", ms_e, n, 2 * (h // 2) * (w // 2))
        report(f"I420 U plane 960x540 extract ({tag})", ms_x, n, (h // 2) * (w // 2))
    ops.set_path(0)
    del frames
    # ---- fused rgb24 1080p frames (reference default flow: U channel of the converted frame)
    h, w, n = 1080, 1920, 512
    # coloured content: three different smooth fields plus noise, so chroma is not a flat 0.5
    frames = torch.stack([planes(n, h, w, seed=s).roll(37 * s, dims=2).roll(11 * s, dims=1) for s in range(3)], dim=3)
    frames[..., 0] = (frames[..., 0].float() * 0.6 + 90).clamp(0, 255).to(torch.uint8)
    frames = frames.contiguous()
    wm, ln = ops.pack_bits(Shuffler(key=0).generate_wm(payload, (1, h * w // 64))[0], device=DEV)
    ms_e = timed(lambda: ops.dwtsvd_embed_rgb8_(frames, wm, ln))
    ms_x = timed(lambda: ops.dwtsvd_extract_rgb8(frames, payload_len=8))
    report("rgb24 fused embed (in place, 6 B/px)", ms_e, n, 6 * h * w)
    report("rgb24 fused extract (3 B/px)", ms_x, n, 3 * h * w)

    # the same kernels on natural-image statistics: the reference's 1080p fixture crop (tests/golden), tiled to
    # 1080p with a little per-frame noise.  Chroma blocks of real pictures are mostly well dominated; the smooth
    # synthetic fields above cross zero in U and V and send most warps through the slow path.
    crop = np.load(os.path.join(ROOT, "tests", "golden", "frame63_crop.npz"))["bgr"]
    reps = (-(-h // crop.shape[0]), -(-w // crop.shape[1]), 1)
    tile = torch.from_numpy(np.tile(crop, reps)[:h, :w].copy()).to(DEV)
    nat = torch.empty((256, h, w, 3), dtype=torch.uint8, device=DEV)
    gnat = torch.Generator(device=DEV).manual_seed(9)
    for f0 in range(0, 256, 32):
        nat[f0:f0 + 32] = (tile[None].float() + 1.5 * torch.randn((32, h, w, 3), device=DEV, generator=gnat)).round().clamp(0, 255).to(torch.uint8)
    ms_e = timed(lambda: ops.dwtsvd_embed_rgb8_(nat, wm, ln))
    ms_x = timed(lambda: ops.dwtsvd_extract_rgb8(nat, payload_len=8))
    report("rgb24 fused embed, natural image tiled (in place, 6 B/px)", ms_e, 256, 6 * h * w)
    report("rgb24 fused extract, natural image tiled (3 B/px)", ms_x, 256, 3 * h * w)
    del nat, tile

    def unfused():
        yuv = ops.bgr8_to_yuv32(frames[:128])
        ops.dwtsvd_embed_(yuv, wm, ln, channel=1)
        return ops.yuv32_to_bgr8(yuv)
    report("rgb24 three-kernel embed (bracket + strided float32 embed + bracket)", timed(unfused), 128, 6 * h * w)
    del frames
    # ---- 8x8 DCT pair on float32 interleaved YUV (the reference layout)
    n = 128
    yuv = torch.rand((n, h, w, 3), device=DEV) * 200 + 20
    ms_m = timed(lambda: ops.dct8_masks(yuv, channel=0))
    masks = ops.dct8_masks(yuv, channel=0)
    ms_e = timed(lambda: ops.dct8_embed_(yuv, masks, wm, ln, alpha=20, channel=1))
    ms_x = timed(lambda: ops.dct8_extract(yuv, masks, alpha=20, payload_len=8, channel=1))
    report("dct8 masks (float32 interleaved, Y read)", ms_m, n, 4 * h * w)
    report("dct8 embed (U read+write)", ms_e, n, 8 * h * w)
    report("dct8 extract (U read)", ms_x, n, 4 * h * w)
    del yuv
    # ---- the same pair on planar uint8 (yuv444 planes): masks from Y, mark in the U plane
    n = 512
    yp = planes(n, h, w, seed=5)
    up = planes(n, h, w, seed=6)
    ms_m = timed(lambda: ops.dct8_masks(yp))
    masks = ops.dct8_masks(yp)
    ms_e = timed(lambda: ops.dct8_embed_(up, masks, wm, ln, alpha=20))
    ms_x = timed(lambda: ops.dct8_extract(up, masks, alpha=20, payload_len=8))
    report("dct8 masks (planar uint8, Y read)", ms_m, n, h * w)
    report("dct8 embed (planar uint8, U read+write)", ms_e, n, 2 * h * w)
    report("dct8 extract (planar uint8, U read)", ms_x, n, h * w)
    del yp, up
    # ---- attack kernels
    y = planes(512, h, w)
    report("jpeg-like requant q75 (in place)", timed(lambda: ops.attack_jpeg_requant_(y, 75)), 512, 2 * h * w)
    small = ops.attack_resize(y, (1280, 720), ops.INTER_AREA)
    report("resize area 1080p -> 720p", timed(lambda: ops.attack_resize(y, (1280, 720), ops.INTER_AREA)), 512, h * w + 1280 * 720)
    report("resize bilinear 720p -> 1080p", timed(lambda: ops.attack_resize(small, (w, h), ops.INTER_LINEAR)), 512, h * w + 1280 * 720)
    del small
    # ---- one read, N marked copies (fingerprinting flow): (1 + N) * W * H bytes per frame
    y = y[:384]
    for copies in (2, 4, 8):
        rows = np.stack([Shuffler(key=0).generate_wm(np.array([int(b) for b in format(c, "08b")]), (1, h * w // 64))[0] for c in range(copies)])
        wmc, lnc = ops.pack_bits(rows, device=DEV)
        out = torch.empty((copies,) + tuple(y.shape), dtype=torch.uint8, device=DEV)
        ms_c = timed(lambda: ops.dwtsvd_embed_copies(y, wmc, lnc, copies, out=out))
        tmp = out[0]
        ms_1 = timed(lambda: [ops.dwtsvd_embed_(y, wmc, lnc, out=tmp, frame_wm_row=None) for _ in range(copies)])
        report(f"{copies} marked copies, one read", ms_c, y.shape[0], (1 + copies) * h * w, single_embeds_ms=round(ms_1, 4))
        del out


if __name__ == "__main__":
    main()
