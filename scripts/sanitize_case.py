"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both kernel families, embed,
extract, votes, DCT pair, fused rgb, attacks on tiny inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))
import numpy as np, torch
from b200wm import ops
from b200wm.vote import SegmentVote

dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
for path in (0, 1):
    ops.set_path(path)
    for (n, h, w) in ((3, 64, 512), (2, 72, 1040), (2, 37, 53)):
        planes = torch.from_numpy(rng.randint(0, 256, (n, h, w)).astype(np.uint8)).to(dev)
        bits = rng.randint(0, 2, (2, max(1, h * w // 64)))
        wm, ln = ops.pack_bits(bits, device=dev)
        rows = torch.tensor([i % 2 for i in range(n)], dtype=torch.int32, device=dev)
        ops.dwtsvd_embed_(planes, wm, ln, frame_wm_row=rows)
        raw, counts = ops.dwtsvd_extract(planes, payload_len=8)
        if h * w // 64 > 0:
            perm = torch.arange(8, dtype=torch.int32, device=dev)
            patterns, packed = ops.vote_finish(counts, h * w // 64, perm)
            SegmentVote(2, 8, dev).add(packed, frame_segment=rows).result()
            ops.vote_counts(raw, h * w // 64, 5)
ops.set_path(0)
yuv = torch.rand((2, 64, 96, 3), device=dev) * 255
wm, ln = ops.pack_bits(rng.randint(0, 2, 64 * 96 // 64), device=dev)
masks = ops.dct8_masks(yuv, channel=0)
ops.dct8_embed_(yuv, masks, wm, ln, channel=1)
ops.dct8_extract(yuv, masks, payload_len=8, channel=1)
ops.dwtsvd_embed_(yuv, wm, ln, channel=1)
ops.dwtsvd_extract(yuv, channel=1, payload_len=8)
rgb = torch.from_numpy(rng.randint(0, 256, (2, 64, 96, 3)).astype(np.uint8)).to(dev)
ops.dwtsvd_embed_rgb8_(rgb, wm, ln)
ops.dwtsvd_extract_rgb8(rgb, payload_len=8)
ops.yuv32_to_bgr8(ops.bgr8_to_yuv32(rgb))
y = torch.from_numpy(rng.randint(0, 256, (2, 64, 96)).astype(np.uint8)).to(dev)
ops.attack_jpeg_requant_(y, 75)
ops.attack_add_noise_(y, torch.randn((2, 64, 96), device=dev))
torch.cuda.synchronize()
print("sanitize case done, launches", ops.kernel_launches())
