#!/usr/bin/env python
"""Where the per-frame plugin flow (Embedder / Extractor on rgb24 numpy frames) spends its host time.
    python scripts/profile_plugin_flow.py [--frames 256]
Prints cProfile's top entries for one warm run of the flow bench.py's e2e_plugin_rgb24 leg times."""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))

import numpy as np      # noqa: E402
import torch            # noqa: E402

from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder      # noqa: E402
from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder    # noqa: E402
from offmark_b200.generator.shuffler import Shuffler                     # noqa: E402
from offmark_b200.degenerator.de_shuffler import DeShuffler              # noqa: E402
from offmark_b200.video.embedder import Embedder                         # noqa: E402
from offmark_b200.video.extractor import Extractor                       # noqa: E402
from offmark_b200.video.memory_io import ArrayReader, BatchReader, BatchWriter      # noqa: E402

PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    args = ap.parse_args()
    n, h, w = args.frames, 1080, 1920
    rng = np.random.default_rng(0)
    clip = rng.integers(16, 236, size=(n, h, w, 3), dtype=np.uint8)

    class Sink:
        def __init__(self):
            self.array = np.zeros((n, h, w, 3), dtype=np.uint8)
            self.k = 0

        def write(self, f):
            self.array[self.k] = f
            self.k += 1

        def close(self):
            pass

    sinks = {}

    def flow():
        enc = DwtDctSvdEncoder()
        enc.read_wm(Shuffler(key=0).generate_wm(PAYLOAD, enc.wm_capacity((h, w, 3))))
        sink = sinks.setdefault(0, None) or sinks.__setitem__(0, Sink()) or sinks[0]
        sink.k = 0
        t0 = time.perf_counter()
        Embedder(ArrayReader(list(clip)), enc, sink, batch_frames=32).start()
        t1 = time.perf_counter()
        ex = Extractor(ArrayReader(list(sink.array)), DwtDctSvdDecoder(), DeShuffler(key=0).set_shape((8,)), batch_frames=32)
        ex.start()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    pinned_clip = torch.from_numpy(clip).pin_memory().numpy()
    writer = BatchWriter(n, (h, w, 3))

    def flow_batch_io():
        enc = DwtDctSvdEncoder()
        enc.read_wm(Shuffler(key=0).generate_wm(PAYLOAD, enc.wm_capacity((h, w, 3))))
        writer.rewind()
        t0 = time.perf_counter()
        Embedder(BatchReader(pinned_clip), enc, writer, batch_frames=32).start()
        t1 = time.perf_counter()
        ex = Extractor(BatchReader(writer.array), DwtDctSvdDecoder(), DeShuffler(key=0).set_shape((8,)), batch_frames=32)
        ex.start()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    for name, fn in (("read()/write() per frame", flow), ("batch protocol, pinned", flow_batch_io)):
        fn()
        for _ in range(4):
            a, b = fn()
            print(f"{name}: embed pass {a * 1e3:.1f} ms, extract pass {b * 1e3:.1f} ms, {n / (a + b):.0f} frames/s")
    t0 = time.perf_counter()
    x = torch.empty(32 * h * w * 3, dtype=torch.uint8, pin_memory=True)
    print(f"pinned allocation of {x.numel() >> 20} MiB: {(time.perf_counter() - t0) * 1e3:.1f} ms")
    src, dst = clip[0], np.empty_like(clip[0])
    t0 = time.perf_counter()
    for _ in range(20):
        np.copyto(dst, src)
    print(f"one-thread frame copy: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms per 1080p rgb24 frame")
    prof = cProfile.Profile()
    prof.enable()
    flow()
    prof.disable()
    stats = pstats.Stats(prof).sort_stats("tottime")
    stats.print_stats(12)
    stats.print_callers("torch.empty")


if __name__ == "__main__":
    main()
