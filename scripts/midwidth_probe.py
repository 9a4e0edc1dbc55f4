import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'video-fingerprinting_b200')); sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import numpy as np, torch
from b200wm import ops
from offmark_b200.generator.shuffler import Shuffler
import bench_extra as be
payload = np.array([0, 1, 1, 0, 0, 1, 0, 1])
for (h, w) in [(720, 1280), (1920, 1080), (1080, 1440), (1152, 1536), (1080, 1664), (1080, 1792)]:
    n = max(64, int(2.0e9 // (h * w)))
    src = be.planes(n, h, w); dst = src.clone()
    wm, ln = ops.pack_bits(Shuffler(key=0).generate_wm(payload, (1, h * w // 64))[0], device=be.DEV)
    out = {}
    for path, tag in ((0, "tma"), (1, "ldg")):
        ops.set_path(path)
        ms_e = be.timed(lambda: ops.dwtsvd_embed_(src, wm, ln, out=dst)); ms_x = be.timed(lambda: ops.dwtsvd_extract(dst, payload_len=8))
        out[tag] = round(n * 3 * h * w / ((ms_e + ms_x) * 1e-3) / 1e9 / be.PEAK, 3)
    print(w, 'x', h, 'tiles_x', w // 8, out, flush=True)
    del src, dst
