"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family and
every launch shape of the TMA path on short planes, checked against nothing but the tool.
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py
Heights are a few tile rows so that the instrumented run stays within a minute or two."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))
import numpy as np, torch
from b200wm import ops
from b200wm.vote import SegmentVote

dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
QUICK = "--quick" in sys.argv


def mark_and_read(planes, n, h, w):
    if h * w // 64 == 0:
        ops.dwtsvd_extract(planes, payload_len=8)
        return
    bits = rng.randint(0, 2, (2, h * w // 64))
    wm, ln = ops.pack_bits(bits, device=dev)
    rows = torch.tensor([i % 2 for i in range(n)], dtype=torch.int32, device=dev)
    out = torch.empty_like(planes) if planes.is_contiguous() else None
    ops.dwtsvd_embed_(planes, wm, ln, frame_wm_row=rows, out=out)        # out of place (tight planes) ...
    ops.dwtsvd_embed_(planes, wm, ln, frame_wm_row=rows)                 # ... and in place
    raw, counts = ops.dwtsvd_extract(planes, payload_len=8)
    perm = torch.arange(8, dtype=torch.int32, device=dev)
    patterns, packed = ops.vote_finish(counts, h * w // 64, perm)
    SegmentVote(2, 8, dev).add(packed, frame_segment=rows).result()
    ops.vote_counts(raw, h * w // 64, 5)


# widths: 64 tiles (smallest TMA strip), 128, 130 (vectorised-load kernels), 192 (three consumer warps), 240 (1080p
# instance), 256 (generic pitch), 480 (4K: column chunks for embed, whole strips of eight warps for extract), 520
# (column chunks for both), 53 (ragged, no tile alignment)
shapes = [(3, 24, 512), (2, 16, 1024), (2, 72, 1040), (2, 24, 1536), (2, 24, 1920), (2, 16, 2048), (1, 16, 3840), (1, 16, 4160), (2, 37, 53)]
if QUICK:
    shapes = [(2, 24, 1536), (2, 24, 1920), (1, 16, 3840), (2, 37, 53)]
for path in (0, 1):
    ops.set_path(path)
    for (n, h, w) in shapes:
        planes = torch.from_numpy(rng.randint(0, 256, (n, h, w)).astype(np.uint8)).to(dev)
        mark_and_read(planes, n, h, w)
ops.set_path(0)
# pitched planes (a window of wider frames) and the three planes of I420 frames (chroma: narrow-plane mode)
wide = torch.from_numpy(rng.randint(0, 256, (2, 24, 2048)).astype(np.uint8)).to(dev)
mark_and_read(wide[:, :, 64:64 + 1920], 2, 24, 1920)
i420 = torch.from_numpy(rng.randint(0, 256, (2, 1920 * 32 * 3 // 2)).astype(np.uint8)).to(dev)
for which, (h, w) in (("y", (32, 1920)), ("u", (16, 960)), ("v", (16, 960))):
    mark_and_read(ops.i420_plane(i420, 32, 1920, which), 2, h, w)
# flat and piecewise-flat tiles (the reference's flat-tile rule has its own code path)
flat = torch.full((2, 24, 1920), 255, dtype=torch.uint8, device=dev)
flat[1, :, 960:] = 30
mark_and_read(flat, 2, 24, 1920)

# float32 interleaved frames: DCT pair (three-kernel and composed forms), DWT/SVD on a chroma channel, copies kernel
yuv = torch.rand((2, 64, 96, 3), device=dev) * 255
wm, ln = ops.pack_bits(rng.randint(0, 2, 64 * 96 // 64), device=dev)
masks = ops.dct8_masks(yuv, channel=0)
ops.dct8_embed_(yuv, masks, wm, ln, channel=1)
ops.dct8_extract(yuv, masks, payload_len=8, channel=1)
ops.dct8_encode_(yuv, yuv, wm, ln, lum_channel=0, channel=1)
ops.dct8_decode(yuv, yuv, payload_len=8, lum_channel=0, channel=1)
ops.dwtsvd_embed_(yuv, wm, ln, channel=1)
ops.dwtsvd_extract(yuv, channel=1, payload_len=8)
ops.dwtsvd_sigma(yuv, channel=1)
# planar uint8 4:4:4: the conversion-free DCT pair
planar = torch.from_numpy(rng.randint(0, 256, (2, 3, 64, 96)).astype(np.uint8)).to(dev)
ops.dct8_encode_(planar[:, 0], planar[:, 1], wm, ln)
ops.dct8_decode(planar[:, 0], planar[:, 1], payload_len=8)
one = torch.from_numpy(rng.randint(0, 256, (1, 64, 96)).astype(np.uint8)).to(dev)
rows2, _ = ops.pack_bits(rng.randint(0, 2, (3, 64 * 96 // 64)), device=dev)
ops.dwtsvd_embed_copies(one, rows2, ln, 3)
# fused rgb24 (aligned and odd widths), colour bracket
for (h, w) in ((64, 96), (40, 136), (24, 1920)):
    rgb = torch.from_numpy(rng.randint(0, 256, (2, h, w, 3)).astype(np.uint8)).to(dev)
    wmr, lnr = ops.pack_bits(rng.randint(0, 2, max(1, h * w // 64)), device=dev)
    ops.dwtsvd_embed_rgb8_(rgb, wmr, lnr)
    ops.dwtsvd_extract_rgb8(rgb, payload_len=8)
    ops.yuv32_to_bgr8(ops.bgr8_to_yuv32(rgb))
# attacks
y = torch.from_numpy(rng.randint(0, 256, (2, 72, 96)).astype(np.uint8)).to(dev)
ops.attack_jpeg_requant_(y, 75)
ops.attack_add_noise_(y, torch.randn((2, 72, 96), device=dev))
ops.attack_resize_roundtrip_(y)
ops.attack_resize(y, (50, 31), ops.INTER_AREA)
ops.attack_resize(y, (131, 97), ops.INTER_LINEAR)
# host-buffer entry points (chunked, three streams)
src = rng.randint(0, 256, (5, 24, 1920)).astype(np.uint8)
dst = np.empty_like(src)
rows_host = ops.pack_bits(rng.randint(0, 2, (2, 24 * 1920 // 64)))[0].cpu().contiguous()
fr = torch.tensor([0, 1, 0, 1, 1], dtype=torch.int32)
perm = np.arange(8, dtype=np.int32)
ops.dwtsvd_mark_host(src, dst, rows_host, frame_wm_row=fr, wm_len=24 * 1920 // 64, chunk_frames=2)
ops.dwtsvd_detect_host(dst, perm, chunk_frames=2)
ops.dwtsvd_mark_verify_host(src, dst, rows_host, perm, frame_wm_row=fr, wm_len=24 * 1920 // 64, chunk_frames=2)
ops.host_scratch_release()
# vote state reset + the publish kernel in a one-rank world (no peers: histogram + ticket only)
v = SegmentVote(3, 8, dev)
v.add(torch.tensor([5, 5, 9, 200], dtype=torch.int64, device=dev), frame_segment=torch.tensor([0, 0, 1, 2], dtype=torch.int32, device=dev))
v.reset()
v.add(torch.tensor([7], dtype=torch.int64, device=dev), frame_segment=torch.tensor([2], dtype=torch.int32, device=dev))
v.result()
torch.cuda.synchronize()
print("sanitize case done, launches", ops.kernel_launches())
