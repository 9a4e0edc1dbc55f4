#!/usr/bin/env python
"""Benchmark of the watermark hot path: 1080p frames/s, embed + extract + vote.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline (BASELINE.json configs[1]): 3000 synthetic 1080p I420 frames per GPU,
resident in HBM; the marked plane is the uint8 Y plane.  One step = one pass of the hot path over the
whole batch:
    embed   (b200wm_dwtsvd_embed,   read Y + write marked Y)                          2*W*H bytes/frame
    extract (b200wm_dwtsvd_extract, read marked Y, raw bits and per-position counts)  W*H bytes/frame
    vote    (b200wm_vote_finish + b200wm_pattern_hist into a persistent per-segment state, then - when N > 1 -
            one all-gather of the 104 KB vote state, enqueued asynchronously so that it overlaps the next
            step's embed; every exchange completes inside the timed region)
Frames are grouped in 2-second segments of 60 frames, every segment carries its own payload (the 8-bit
segment number, tests/segment_mark_detect_hls.py:42-55 in the reference).  Under torchrun (N > 1) every
rank owns its own 3000 frames (weak scaling).  Rank 0 prints ONE JSON line.

Besides the headline the line carries legs that are measured OUTSIDE the headline's timed region:
    parity               8 frames of the batch against the oracle (embed LSB distance, raw-bit agreement, votes)
    e2e                  the same metric through b200wm_dwtsvd_mark_verify_host on pinned HOST planes
    e2e_mark_then_detect ... through b200wm_dwtsvd_mark_host + b200wm_dwtsvd_detect_host (round-1 e2e: 3 trips)
    e2e_plugin_rgb24     ... through the plugin objects Embedder / Extractor on rgb24 host frames (per-frame read / write;
                         e2e_plugin_rgb24_batch_io: the same objects with the optional batch reader / writer protocol)
    config3_4k           BASELINE configs[2]: 4K planes, 256 frames per GPU, with its own roofline (every N)
    pair_dct8            the 8x8-DCT plugin pair (DctEncoder / DctDecoder) on 4:4:4 planar uint8 (N = 1)
    attacks              BASELINE configs[4]: BER after JPEG-like / noise / resize on 10,000 frames, against the
                         oracle extractor on 64 of the same attacked frames per attack (N = 1)
    cpu_baseline         the oracle's per-block port (the reference's cost structure) on all host cores (N = 1)

`--impl reference` times the reference's CPU implementation of the same path (the oracle's per-block port of its
Python loops: the reference is pure Python and cannot travel to the GPU box) on all host cores, on a bounded
sample of the very frames the GPU arm marks (same generator, same seed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))

import numpy as np          # noqa: E402
import torch                # noqa: E402

H, W = 1080, 1920
SEGMENT_FRAMES = 60
PAYLOAD_LEN = 8
KEY = 0
SEED = 20261018
METRIC = "frames_per_sec_1080p_embed_extract"
PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--frames", type=int, default=3000, help="frames per GPU")
    p.add_argument("--e2e-frames", type=int, default=3000)
    p.add_argument("--e2e-chunk", type=int, default=0, help="frames per chunk inside the host-buffer calls (0: the library's 64 MB default)")
    p.add_argument("--plugin-frames", type=int, default=256)
    p.add_argument("--attack-frames", type=int, default=10000)
    p.add_argument("--attack-oracle-frames", type=int, default=64)
    p.add_argument("--cpu-frames-per-core", type=int, default=8)     # ~2 s per frame and core: 10-20 s of CPU work
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the parity / 4K / dct8 / attack legs")
    p.add_argument("--size", default="1080p", choices=["1080p", "4k"],
                   help="1080p: BASELINE configs[1] (the headline); 4k: configs[2] as the headline (750 frames per GPU)")
    p.add_argument("--path", type=int, default=0, choices=[0, 1],
                   help="0: TMA-staged persistent kernels (default), 1: vectorised-load kernels (A/B evidence)")
    return p.parse_args()


# --------------------------------------------------------------------------------- synthetic data
def generate_i420(n_frames, device, rank, h=None, w=None):
    """[n_frames, w*h*3/2] uint8 I420 frames generated on the device.  Generated in
    chunks of 50 frames from one seeded generator, so the first frames do not depend on how many follow."""
    h, w = h or H, w or W
    fb = w * h * 3 // 2
    buf = torch.empty((n_frames, fb), dtype=torch.uint8, device=device)
    gen = torch.Generator(device=device).manual_seed(SEED + rank)
    xx = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    yy = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
    chunk = 50 if h <= 1080 else 10
    for f0 in range(0, n_frames, chunk):
        n = min(chunk, n_frames - f0)
        f = (torch.arange(f0, f0 + n, device=device, dtype=torch.float32) + rank * n_frames)[:, None, None]
        y = 128 + 80 * torch.sin(2 * np.pi * (3 * xx / w + f / 97)) * torch.cos(2 * np.pi * (2 * yy / h + f / 53))
        y = y + 6.0 * torch.randn((n, h, w), device=device, generator=gen)
        buf[f0:f0 + n, :w * h] = y.round().clamp(16, 235).to(torch.uint8).reshape(n, -1)
        uv = 128 + 4.0 * torch.randn((n, w * h // 2), device=device, generator=gen)
        buf[f0:f0 + n, w * h:] = uv.round().clamp(16, 240).to(torch.uint8)
    return buf


def y_planes(i420, h=None, w=None):
    """[N, h, w] view of the Y planes inside the I420 buffer (frame stride w*h*3/2)."""
    h, w = h or H, w or W
    return i420.as_strided((i420.shape[0], h, w), (w * h * 3 // 2, w, 1))


def segment_rows(first_segment, n_segments, h=None, w=None):
    """Watermark rows (one per segment) and the payloads they carry."""
    from offmark_b200.generator.shuffler import Shuffler
    h, w = h or H, w or W
    # 8-bit segment numbers, skipping 0 and 255: an all-equal payload defeats the reference's
    # adaptive threshold 0.5*(max+min) (de_shuffler.py:20) by construction
    payloads = [np.array([int(b) for b in format(1 + (first_segment + s) % 254, "08b")]) for s in range(n_segments)]
    rows = np.stack([Shuffler(key=KEY).generate_wm(p, (1, h * w // 64))[0] for p in payloads])
    return rows, np.stack(payloads)


def load_peak():
    peaks = {}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peaks = json.load(f)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, ("MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)")


def timed_ms(fn, warmup=2, steps=8):
    """Average milliseconds of ``fn`` on the current stream (CUDA events, after warm-up, synchronised both sides)."""
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


def max_over_ranks(value, dev, world):
    if world <= 1:
        return float(value)
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML in a thread, 5 ms period;
    nvidia-smi as a fallback).  The device index is the physical one behind CUDA_VISIBLE_DEVICES."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.samples, self.stop_flag, self.thread, self.nvml, self.smi = [], False, None, None, None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                index = int(vis.split(",")[index])
            except (ValueError, IndexError):
                pass
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            self._start_smi(index)

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.samples.append((time.time(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def _start_smi(self, index):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                         "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.smi = None

    def _read_smi(self):
        for line in self.smi.stdout:
            parts = [x.strip() for x in line.split(",")]
            try:
                mask = sum(1 << i for i, v in enumerate(parts[2:6]) if v.lower().startswith("active"))
                self.samples.append((time.time(), float(parts[0]), mask, float(parts[1])))
            except (ValueError, IndexError):
                continue

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.smi is not None:
            time.sleep(0.06)
            self.smi.terminate()
        if self.nvml is None and self.smi is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        used = inside or self.samples[-2:]
        reasons = set()
        for s in used:
            if self.nvml is not None:
                for name, attr in self.REASONS:
                    if s[2] & getattr(self.nvml, attr):
                        reasons.add(name)
            else:
                for i, (name, _) in enumerate(self.REASONS):
                    if s[2] & (1 << i):
                        reasons.add(name)
        max_mhz = self.max_mhz if self.nvml is not None else (max(s[3] for s in used) if used else None)
        return {"sm_mhz": float(np.median([s[1] for s in used])) if used else None, "sm_max_mhz": max_mhz,
                "reasons": sorted(reasons), "samples": len(inside), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------- CPU legs (oracle; bench.py may time / check with it)
def _single_thread_env():
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"              # one process per core; inherited by the spawned workers


def _cpu_worker(args):
    planes, wm_row = args
    from oracle import dwt_dct_svd as o_svd, payload as o_pay
    out = []
    for y in planes:
        yuv = np.zeros(y.shape + (3,), dtype=np.float32)
        yuv[:, :, 1] = y
        o_svd.encode_per_block(yuv, wm_row[None, :])
        yuv[:, :, 1] = np.around(np.clip(yuv[:, :, 1], 0, 255))
        bits = o_svd.decode_per_block(yuv)
        out.append(o_pay.degenerate(bits, PAYLOAD_LEN, KEY))
    return out


def _cpu_worker_vectorised(args):
    planes, wm_row = args
    from oracle import dwt_dct_svd as o_svd, payload as o_pay
    return [o_pay.degenerate(o_svd.extract_plane(o_svd.embed_plane_u8(y, wm_row)), PAYLOAD_LEN, KEY) for y in planes]


def _cpu_worker_dct8(args):
    """The 8x8-DCT pair as the reference runs it (per-block cv2.dct loops for masks, embed and extract)."""
    ys, us, wm_row = args
    from oracle import dct8 as o_dct, payload as o_pay
    out = []
    for y, u in zip(ys, us):
        yuv = np.zeros(y.shape + (3,), dtype=np.float32)
        yuv[:, :, 0], yuv[:, :, 1] = y, u
        o_dct.encode(yuv, wm_row[None, :])
        yuv[:, :, 1] = np.around(np.clip(yuv[:, :, 1], 0, 255))
        out.append(o_pay.degenerate(o_dct.decode(yuv), PAYLOAD_LEN, KEY))
    return out


def _parity_worker(args):
    """One frame of the batch against the oracle: marked plane vs oracle embed, raw bits of the marked plane vs
    oracle extract, voted payload vs oracle vote.  Blocks on a quantisation boundary (oracle/knife_edge.py) are
    counted and reported, never silently dropped."""
    src, marked, bits_gpu, pattern_gpu, wm_row = args
    from oracle import dwt_dct_svd as o_svd, payload as o_pay, knife_edge as ke
    want = o_svd.embed_plane_u8(src, wm_row)
    d = np.abs(marked.astype(np.int16) - want.astype(np.int16))
    _, floor_src, _ = ke.knife_edge_blocks(src.astype(np.float32))
    ok = ~ke.tile_mask_to_pixels(floor_src, src.shape)
    bits_ref = o_svd.extract_plane(marked)[0].astype(np.uint8)
    edge_marked, _, _ = ke.knife_edge_blocks(marked.astype(np.float32))
    mism = bits_gpu[:bits_ref.size] != bits_ref
    on_edge = np.zeros(bits_ref.size, dtype=bool)
    on_edge[:edge_marked.size] = edge_marked
    return {"max_abs": int(d[ok].max(initial=0)), "max_abs_all": int(d.max(initial=0)), "px_diff": int((d > 0).sum()), "px": int(d.size),
            "bits": int(bits_ref.size), "mismatch": int(mism.sum()), "mismatch_off_boundary": int((mism & ~on_edge).sum()),
            "masked_embed_blocks": int(floor_src.sum()), "masked_extract_blocks": int(edge_marked.sum()),
            "blocks": int(edge_marked.size),
            "payload_same": bool(np.array_equal(o_pay.degenerate(bits_ref.reshape(1, -1), PAYLOAD_LEN, KEY), pattern_gpu))}


def _attack_worker(args):
    """Reference extractor (oracle) on one attacked frame: raw BER vs the embedded bits, agreement with the GPU
    extractor's raw bits on and off the quantisation boundaries, identical vote."""
    plane, bits_gpu, pattern_gpu, wm_row = args
    from oracle import dwt_dct_svd as o_svd, payload as o_pay, knife_edge as ke
    ref = o_svd.extract_plane(plane)[0].astype(np.uint8)
    edge, _, _ = ke.knife_edge_blocks(plane.astype(np.float32))
    mism = bits_gpu[:ref.size] != ref
    on_edge = np.zeros(ref.size, dtype=bool)
    on_edge[:edge.size] = edge
    return (float((ref != wm_row[:ref.size]).mean()), int(mism.sum()), int((mism & ~on_edge).sum()), int(edge.sum()), int(ref.size),
            bool(np.array_equal(o_pay.degenerate(ref.reshape(1, -1), PAYLOAD_LEN, KEY), pattern_gpu)))


class CpuPool:
    """Spawned worker processes (one per host core), started once and shared by the CPU legs."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        _single_thread_env()
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_cpu_worker, [(np.zeros((0, 8, 8), dtype=np.uint8), np.zeros(1, dtype=np.int64))] * self.cores)   # imports

    def map(self, fn, jobs):
        return self.pool.map(fn, jobs, chunksize=1)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_vectorised(pool, planes_host, wm_row, frames_per_core=2):
    """A 'fair CPU' line next to the reference-shaped baseline: the oracle's VECTORISED form (one batched
    np.linalg.svd over all 32,400 blocks of a plane) on every host core.  NOT the reference's cost structure."""
    cores = min(pool.cores, len(planes_host))
    jobs = [(planes_host[i::cores][:frames_per_core], wm_row) for i in range(cores)]
    n = sum(len(j[0]) for j in jobs)
    t0 = time.perf_counter()
    pool.map(_cpu_worker_vectorised, jobs)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port (vectorised numpy, not the reference's per-block loop)",
            "sample": f"{n} 1080p frames, embed+extract+per-frame vote, batched SVD over all blocks of a plane, {cores} processes, {dt:.1f} s"}


def cpu_baseline(pool, planes_host, wm_row, frames_per_core=1, band_rows=None):
    """fps of the reference's per-block CPU path (oracle port) using every host core.  Each core gets
    ``frames_per_core`` bands of ``band_rows`` rows (a whole frame by default; blocks are independent, so a band
    of whole tile rows is a valid bounded sample)."""
    h = planes_host.shape[1]
    band_rows = max(8, min(h, (band_rows or h) // 8 * 8))
    cores = min(pool.cores, len(planes_host))
    jobs = [(planes_host[i::cores][:frames_per_core, :band_rows], wm_row) for i in range(cores)]
    n = sum(len(j[0]) for j in jobs)
    t0 = time.perf_counter()
    pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    frames = n * band_rows / h
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"{n} bands of {band_rows}x{planes_host.shape[2]} ({frames:.2f} frames) of the batch the GPU arm marks, "
                      f"embed+extract+per-frame vote, oracle per-block port (cv2.dct + np.linalg.svd per 4x4 block, like the "
                      f"reference), {cores} processes, {dt:.1f} s"}


def reference_frames(n):
    """The first ``n`` Y planes of rank 0's batch - the frames the GPU arm marks - regenerated with the same seeded
    device generator when a GPU is visible (the reference arm is its own process), else the numpy recipe."""
    if torch.cuda.is_available():
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        total = (n + 49) // 50 * 50
        planes = y_planes(generate_i420(total, dev, 0))[:n].cpu().numpy()
        return np.ascontiguousarray(planes), "first frames of rank 0's batch (same device generator and seed as the GPU arm)"
    from oracle import synth
    return np.stack([synth.luma_plane_u8(H, W, f, SEED) for f in range(n)]), "numpy recipe of SURVEY §8d config 2 (no GPU visible)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool = CpuPool()
    cores = pool.cores
    planes, origin = reference_frames(cores)
    rows, _ = segment_rows(0, 1)
    wm_row = rows[0]
    # size the per-step sample so that the whole run stays near two minutes: one core needs about
    # 2 s per 1080p frame, i.e. ~15 ms per tile row
    probe = cpu_baseline(pool, planes, wm_row, 1, band_rows=64)
    sec_per_row = probe["seconds"] / 64
    budget = 120.0 / max(1, args.steps + args.warmup)
    band_rows = int(max(8, min(H, budget / sec_per_row)))
    results = []
    for step in range(args.warmup + args.steps):
        res = cpu_baseline(pool, planes, wm_row, 1, band_rows=band_rows)
        if step >= args.warmup:
            results.append(res)
    pool.close()
    fps = float(np.mean([r["value"] for r in results]))
    last = results[-1]
    frames = cores * (max(8, band_rows // 8 * 8)) / H
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * frames / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.frames),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": last["cores"], "kind": "port", "sample": last["sample"],
                             "frames": origin},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(frames_per_gpu, h=None, w=None):
    h, w = h or H, w or W
    name = ("4K (3840x2160) YUV420 synthetic frames sharded across the GPUs, DwtDctSvd embed + extract + vote on the Y plane"
            if h != 1080 else "1080p30 YUV420 synthetic batch, DwtDctSvd embed + extract + vote on the Y plane")
    return {"workload": name, "frames_per_gpu": frames_per_gpu, "height": h, "width": w, "io_dtype": "u8",
            "layout": "planar I420 in HBM", "payload_bits": PAYLOAD_LEN, "segment_frames": SEGMENT_FRAMES, "scale": 15, "blk": 4,
            "cache": f"inputs larger than L2 ({frames_per_gpu * h * w / 1e9:.1f} GB of Y planes per pass vs 126 MB L2)"}


# --------------------------------------------------------------------------------- B200 arm
class Batch:
    """One GPU's resident batch and everything a step needs."""

    def __init__(self, ops, dev, rank, world, n_frames, h, w):
        from b200wm.vote import SegmentVote
        from offmark_b200.degenerator.de_shuffler import DeShuffler
        self.ops, self.dev, self.rank, self.world, self.n, self.h, self.w = ops, dev, rank, world, n_frames, h, w
        self.n_seg_local = (n_frames + SEGMENT_FRAMES - 1) // SEGMENT_FRAMES
        self.i420 = generate_i420(n_frames, dev, rank, h, w)
        self.src = y_planes(self.i420, h, w)
        self.marked_i420 = self.i420.clone()
        self.dst = y_planes(self.marked_i420, h, w)
        self.rows, self.payloads = segment_rows(rank * self.n_seg_local, self.n_seg_local, h, w)
        self.wm_packed, self.wm_len = ops.pack_bits(self.rows, device=dev)
        self.frame_row = (torch.arange(n_frames, device=dev, dtype=torch.int32) // SEGMENT_FRAMES).contiguous()
        self.frame_seg = (self.frame_row + rank * self.n_seg_local).contiguous()
        self.block_num, _, self.words = ops.geometry(h, w)
        self.raw_bits = torch.empty((n_frames, self.words), dtype=torch.int32, device=dev)
        self.pos_counts = torch.empty((n_frames, PAYLOAD_LEN), dtype=torch.int32, device=dev)
        self.deg = DeShuffler(key=KEY).set_shape((PAYLOAD_LEN,))
        # two persistent vote states: the asynchronous exchange of step i overlaps step i+1, which fills the other one
        owned = (rank * self.n_seg_local, self.n_seg_local)
        self.exchange = "none (one GPU)"
        self.votes = None
        if world > 1 and os.environ.get("B200WM_EXCHANGE", "nvlink") == "nvlink":
            try:        # peer-mapped state: the histogram kernel stores its finished block into every peer's buffer itself
                self.votes = [SegmentVote(self.n_seg_local * world, PAYLOAD_LEN, dev, owned=owned, symmetric=True) for _ in range(2)]
                self.exchange = "fused into the histogram kernel: 128-bit stores of the rank's block into every peer's buffer over NVLink, flag per rank (b200wm_pattern_hist_publish)"
            except Exception as exc:        # no peer mapping on this box: the NCCL path below
                self.votes = None
                self.exchange_note = f"symmetric memory unavailable ({type(exc).__name__}: {str(exc)[:120]})"
        if self.votes is None:
            self.votes = [SegmentVote(self.n_seg_local * world, PAYLOAD_LEN, dev, owned=owned) for _ in range(2)]
            if world > 1:
                self.exchange = ("one asynchronous NCCL all-gather of the vote state per step, overlapped with the next step's embed; "
                                 "the last one completes inside the timed region")
        self.step_no = 0
        self.patterns = None

    def step(self, events=None):
        ops = self.ops
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if events is not None else None
        if e: e[0].record()
        ops.dwtsvd_embed_(self.src, self.wm_packed, self.wm_len, scale=15.0, frame_wm_row=self.frame_row, out=self.dst)
        if e: e[1].record()
        ops.dwtsvd_extract(self.dst, scale=15.0, payload_len=PAYLOAD_LEN, raw_bits=self.raw_bits, pos_counts=self.pos_counts)
        if e: e[2].record()
        self.patterns, packed = self.deg.degenerate_counts(self.pos_counts, self.block_num)
        vote = self.votes[self.step_no & 1]
        vote.reset()                        # joins the exchange this state started two steps ago, then one launch
        vote.add(packed, frame_segment=self.frame_seg, order_offset=self.rank * self.n)
        vote.combine(async_op=True)         # no-op on one GPU
        if e:
            e[3].record()
            events.append(e)
        self.step_no += 1
        return vote

    def finish(self):
        for v in self.votes:
            v.wait()


def run_headline(args, ops, batch, dev, local_rank, world):
    import torch.distributed as dist

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        batch.step()
    batch.finish()
    fence()
    launches0 = ops.kernel_launches()
    marks = []
    t_wall0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        vote = batch.step(marks)
    batch.finish()                           # the last exchange completes inside the timed region
    stop.record()
    fence()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = ops.kernel_launches() - launches0
    elapsed_ms = max_over_ranks(start.elapsed_time(stop), dev, world)
    return vote, marks, elapsed_ms, launches, clocks


def main():
    global H, W, METRIC
    args = parse_args()
    if args.size == "4k":                 # BASELINE configs[2] as the headline; the default line stays the 1080p one
        H, W = 2160, 3840
        METRIC = "frames_per_sec_4k_embed_extract"
        if args.frames == 3000:
            args.frames = 750
        args.e2e_frames = min(args.e2e_frames, args.frames)
        args.no_cpu_baseline = args.no_extra = True
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from b200wm import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the watermark kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ops.set_path(args.path)
    n_frames = args.frames
    batch = Batch(ops, dev, rank, world, n_frames, H, W)
    vote, marks, elapsed_ms, launches, clocks = run_headline(args, ops, batch, dev, local_rank, world)

    # ---- accuracy of what the timed steps produced (outside the timed region)
    result = vote.result()
    seg_ok = 0
    for s in range(batch.n_seg_local):
        pattern, freq, _, _ = result[rank * batch.n_seg_local + s]
        seg_ok += int(pattern is not None and np.array_equal(pattern, batch.payloads[s]))
    freqs = [result[rank * batch.n_seg_local + s][1] for s in range(batch.n_seg_local)]
    votes_ok = all(np.array_equal(result[rank * batch.n_seg_local + s][2], batch.payloads[s] * result[rank * batch.n_seg_local + s][3])
                   for s in range(batch.n_seg_local))
    bits = ops.unpack_bits(batch.raw_bits[:64], batch.block_num)
    raw_acc = float((bits == batch.rows[(np.arange(len(bits)) // SEGMENT_FRAMES)]).mean())
    frame_ok = float((batch.patterns.cpu().numpy() == batch.payloads[np.arange(n_frames) // SEGMENT_FRAMES]).all(axis=1).mean())

    # ---- per-kernel times
    k_embed = float(np.mean([e[0].elapsed_time(e[1]) for e in marks]))
    k_extract = float(np.mean([e[1].elapsed_time(e[2]) for e in marks]))
    k_vote = float(np.mean([e[2].elapsed_time(e[3]) for e in marks]))
    ms_per_step = elapsed_ms / args.steps
    value = n_frames * world / (ms_per_step / 1000.0)

    peak, peak_src = load_peak()
    embed_gbs = 2.0 * W * H * n_frames / (k_embed * 1e-3) / 1e9
    extract_gbs = 1.0 * W * H * n_frames / (k_extract * 1e-3) / 1e9
    suffix = "_tma_kernel" if args.path == 0 else "_kernel"
    dominant = ("dwtsvd_embed" if k_embed >= k_extract else "dwtsvd_extract") + suffix
    achieved = embed_gbs if k_embed >= k_extract else extract_gbs
    step_gbs = 3.0 * W * H * n_frames / (ms_per_step * 1e-3) / 1e9

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this very launch size
    # (profiles/r02_traffic.json, else round 1's); null when the workload differs from the captured one
    traffic = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if traffic is None and os.path.exists(tpath) and n_frames == 3000 and H == 1080:
            with open(tpath) as f:
                cap = json.load(f)["kernels"].get(dominant)
            if cap and cap.get("frames") == n_frames:
                traffic = cap["traffic_bytes"]

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n_frames),
        "gpu_launches": launches, "clocks": clocks,
        "kernel_path": "tma_persistent" if args.path == 0 else "ldg_vectorised",
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": (2 if dominant.startswith("dwtsvd_embed") else 1) * W * H * n_frames,
                     "frac_of_nominal_8000_GBs": achieved / 8000.0, "step_frac_of_nominal_8000_GBs": step_gbs / 8000.0},
        "kernels": {"embed_ms": k_embed, "embed_GBs": embed_gbs, "extract_ms": k_extract, "extract_GBs": extract_gbs,
                    "vote_ms": k_vote, "step_GBs": step_gbs, "step_frac_of_peak": step_gbs / peak,
                    "unaccounted_ms_per_step": ms_per_step - k_embed - k_extract - k_vote,
                    "combine": batch.exchange, "combine_note": getattr(batch, "exchange_note", None),
                    "roofline_fps_per_gpu": peak * 1e9 / (3.0 * W * H)},
        "config4_hls_segments": {"config": "BASELINE configs[3]: per-segment payload bits (8-bit segment number, 60-frame segments), Counter "
                                           "vote per segment, state exchanged across the GPUs", "segments_total": batch.n_seg_local * world,
                                 "segments_on_this_rank": batch.n_seg_local, "most_common_pattern_is_payload": seg_ok / batch.n_seg_local,
                                 "min_frequency": float(min(f for f in freqs if f is not None)),
                                 "per_bit_votes_equal_payload_times_frames": bool(votes_ok)},
        "bit_accuracy": {"note": "against the EMBEDDED ground truth; agreement with the reference is in `parity`",
                         "segments_exact": seg_ok / batch.n_seg_local, "frames_exact": frame_ok, "raw_bits_first_64_frames": raw_acc},
    }

    pool = None
    want_pool = rank == 0 and (not args.no_extra or (world == 1 and not args.no_cpu_baseline))
    if want_pool:
        pool = CpuPool()
    if rank == 0 and not args.no_extra:
        line["parity"] = parity_block(pool, batch, ops)
    if not args.no_e2e:
        e2e = run_e2e(args, ops, batch, dev, world)
        line.update(e2e)
    if not args.no_extra and H == 1080:
        line["config3_4k"] = leg_4k(ops, dev, rank, world, peak)
        if world == 1:
            line["pair_dct8"] = leg_dct8(ops, batch, dev, peak, pool)
            line["attacks"] = leg_attacks(args, ops, batch, dev, pool)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = pool.cores
        sample = batch.src[:cores * args.cpu_frames_per_core].cpu().numpy()
        line["cpu_baseline"] = cpu_baseline(pool, sample, batch.rows[0], args.cpu_frames_per_core)
        line["cpu_vectorised"] = cpu_vectorised(pool, sample, batch.rows[0], 2)
    if pool is not None:
        pool.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------- parity of the benchmark batch itself
def parity_block(pool, batch, ops, k=8):
    """K frames of the batch the timed steps just processed (source and marked planes downloaded from HBM, raw bits
    and voted patterns as the kernels left them) against the oracle's embedder / extractor / vote."""
    n = batch.n
    idx = sorted({0, min(59, n - 1), min(60, n - 1), n // 2 - 1, n // 2, (2 * n) // 3, max(0, n - 60), n - 1})[:k]
    sel = torch.tensor(idx, device=batch.dev)
    src = batch.src[sel].cpu().numpy()
    marked = batch.dst[sel].cpu().numpy()
    bits = ops.unpack_bits(batch.raw_bits[sel], batch.block_num)
    patterns = batch.patterns[sel].cpu().numpy()
    rows = batch.rows[np.array(idx) // SEGMENT_FRAMES]
    res = pool.map(_parity_worker, [(src[i], marked[i], bits[i], patterns[i], rows[i]) for i in range(len(idx))])
    total_bits = sum(r["bits"] for r in res)
    blocks = sum(r["blocks"] for r in res)
    return {"frames_checked": len(idx), "frame_indices": idx, "checker": "oracle (vectorised restatement, pinned to the reference by oracle/make_golden.py)",
            "embed_max_abs_lsb": max(r["max_abs"] for r in res),
            "embed_max_abs_lsb_incl_knife_edge_blocks": max(r["max_abs_all"] for r in res),
            "embed_frac_pixels_diff": sum(r["px_diff"] for r in res) / sum(r["px"] for r in res),
            "raw_bit_agreement": 1.0 - sum(r["mismatch"] for r in res) / total_bits,
            "raw_bit_mismatches": sum(r["mismatch"] for r in res),
            "raw_bit_mismatches_off_boundary": sum(r["mismatch_off_boundary"] for r in res),
            "voted_payload_identical": all(r["payload_same"] for r in res),
            "knife_edge_blocks_masked": {"embed_floor": sum(r["masked_embed_blocks"] for r in res),
                                         "extract_bit": sum(r["masked_extract_blocks"] for r in res), "of_blocks": blocks,
                                         "rule": "sigma_0 within 2^-21 (relative) of k*scale or (k+1/2)*scale, flat tiles never masked"}}


# --------------------------------------------------------------------------------- end-to-end legs (host buffers)
def gpu_cpu_affinity(index):
    """CPUs on the NUMA node of GPU ``index`` (NVML), or None.  Pinned staging buffers are placed on the
    node of the thread that allocates them; from the far socket the PCIe copies of the end-to-end path
    measured a third slower, so the bench binds itself next to its GPU while it allocates and streams."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def run_e2e(args, ops, batch, dev, world):
    before = os.sched_getaffinity(0)
    near = gpu_cpu_affinity(dev.index or 0)
    if near:
        os.sched_setaffinity(0, near)
    try:
        res = _run_e2e(args, ops, batch, dev, world)
    finally:
        os.sched_setaffinity(0, before)
    res["e2e"]["host_affinity"] = f"{len(near)} CPUs of the GPU's NUMA node" if near else "unbound"
    return res


def _wall_fps(fn, n, world, dev, steps):
    """Wall-clock frames/s of ``fn`` (a synchronous host-buffer call), barrier on both sides, max over ranks."""
    import torch.distributed as dist
    fn()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    dt = max_over_ranks(time.perf_counter() - t0, dev, world)
    return n * world * steps / dt, dt, out


def _run_e2e(args, ops, batch, dev, world):
    """The same metric through the reference-facing calls with HOST buffers; every copy is inside the timed region.

    e2e                   b200wm_dwtsvd_mark_verify_host: upload source Y, embed, extract + vote on the resident marked
                          chunk, download marked Y + patterns (the reference's mark-then-verify step): 2*W*H on the link
    e2e_mark_then_detect  b200wm_dwtsvd_mark_host, then b200wm_dwtsvd_detect_host on the marked host planes: 3*W*H
    e2e_plugin_rgb24      Embedder(batch_frames=32).start() then Extractor(batch_frames=32).start() on rgb24 host frames
                          (pageable numpy arrays, as FileDecoder yields them): the plugin objects' own flow, 9 B/pixel
    """
    n = min(args.e2e_frames, batch.n)
    h, w = batch.h, batch.w
    host_in = torch.empty((n, h, w), dtype=torch.uint8, pin_memory=True)
    host_marked = torch.empty((n, h, w), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(batch.src[:n])
    torch.cuda.synchronize()
    rows_host = batch.wm_packed.cpu().contiguous()                   # packed payload rows [segments, words] on the host
    frame_row_host = batch.frame_row[:n].cpu()
    perm = batch.deg.payload_idx
    want = batch.payloads[np.arange(n) // SEGMENT_FRAMES]
    steps = max(2, min(args.steps, 3))

    def mark_verify():
        return ops.dwtsvd_mark_verify_host(host_in, host_marked, rows_host, perm, scale=15.0, frame_wm_row=frame_row_host,
                                           wm_len=batch.wm_len, chunk_frames=args.e2e_chunk)

    def mark_then_detect():
        ops.dwtsvd_mark_host(host_in, host_marked, rows_host, scale=15.0, frame_wm_row=frame_row_host, wm_len=batch.wm_len)
        return ops.dwtsvd_detect_host(host_marked, perm, scale=15.0)

    fps1, dt1, pat1 = _wall_fps(mark_verify, n, world, dev, steps)
    marked_equal = bool(torch.equal(host_marked[:64], batch.dst[:64].cpu()))       # the bytes the device-resident step wrote
    fps2, dt2, pat2 = _wall_fps(mark_then_detect, n, world, dev, steps)

    # what the link itself delivers on this box (plain pinned copies, one direction at a time and both together)
    m = min(n, 500)
    scratch = torch.empty((m, h, w), dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(device=dev)

    def timed_copy(up, down):
        torch.cuda.synchronize()
        if world > 1:                   # every rank copies at the same time: the probe sees the box's shared host side
            import torch.distributed as dist
            dist.barrier()
        t0 = time.perf_counter()
        if up:
            scratch.copy_(host_in[:m], non_blocking=True)
        if down:
            with torch.cuda.stream(side):
                host_marked[:m].copy_(batch.src[:m], non_blocking=True)
        torch.cuda.synchronize()
        return m * h * w / (time.perf_counter() - t0) / 1e9
    timed_copy(True, True)
    link = {"h2d_GBs": round(timed_copy(True, False), 1), "d2h_GBs": round(timed_copy(False, True), 1),
            "each_way_GBs_when_both": round(timed_copy(True, True), 1),
            "note": "plain pinned copies of 500 frames on rank 0, all ranks copying at the same time (barrier before each)"}
    del scratch
    out = {
        "e2e": {"value": fps1, "unit": "frames/s", "h2d_bytes_per_step": n * h * w, "d2h_bytes_per_step": n * h * w + n * PAYLOAD_LEN,
                "frames_per_gpu": n, "steps": steps, "each_way_GBs_per_gpu": n * h * w * steps / dt1 / 1e9, "link_probe": link,
                "path": "b200wm_dwtsvd_mark_verify_host on pinned host Y planes: H2D source, embed, extract + vote on the resident "
                        "marked chunk, D2H marked planes + patterns; 32-frame chunks over 3 buffers and 3 streams inside the library",
                "frames_exact": float((pat1 == want).all(axis=1).mean()), "marked_equals_device_path": marked_equal},
        "e2e_mark_then_detect": {"value": fps2, "unit": "frames/s", "h2d_bytes_per_step": 2 * n * h * w,
                                 "d2h_bytes_per_step": n * h * w + n * PAYLOAD_LEN, "frames_per_gpu": n, "steps": steps,
                                 "path": "b200wm_dwtsvd_mark_host + b200wm_dwtsvd_detect_host (round 1's e2e: the marked planes "
                                         "cross the link twice)", "frames_exact": float((pat2 == want).all(axis=1).mean())},
    }
    del host_in, host_marked
    if h == 1080:
        out["e2e_plugin_rgb24"] = plugin_leg(args, ops, batch, dev, world)
        out["e2e_plugin_rgb24_batch_io"] = plugin_leg(args, ops, batch, dev, world, pinned_io=True)
    return out


def plugin_leg(args, ops, batch, dev, world, pinned_io=False):
    """The reference's own driver flow on its own data format: rgb24 host frames -> Embedder.start() -> marked rgb24
    host frames -> Extractor.start() -> per-frame patterns (video/embedder.py:17-39, video/extractor.py:17-34), with
    the drop-in classes in batched mode.  Python-level frame handling is inside the timed region."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.extractor import Extractor
    from offmark_b200.video.memory_io import ArrayReader, BatchReader, BatchWriter

    n, h, w = min(args.plugin_frames, batch.n), batch.h, batch.w
    # rgb24 frames whose luma is the batch's Y plane and whose chroma is mildly coloured (so U is not a flat 0.5)
    y = batch.src[:n].float()
    cb = y_planes(batch.i420, h, w)[:n].roll(31, dims=2).float() * 0.25 + 96.0
    rgb = torch.stack([(y * 0.8 + cb * 0.2 + 10).clamp(0, 255), y, (y * 0.7 + 60 - cb * 0.1).clamp(0, 255)], dim=3).round().to(torch.uint8)
    clip = rgb.cpu().pin_memory().numpy() if pinned_io else rgb.cpu().numpy()
    del rgb, y, cb

    class Sink:                                          # consumes a frame inside write(), like the ffmpeg pipe
        def __init__(self):
            self.array = np.zeros((n, h, w, 3), dtype=np.uint8)      # touched once: write() is a warm copy
            self.k = 0

        def write(self, f):
            self.array[self.k] = f
            self.k += 1

        def close(self):
            pass

    def flow():
        enc = DwtDctSvdEncoder()
        enc.read_wm(Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity((h, w, 3))))
        if pinned_io:       # the optional batch protocol: batches move from / into the reader's and writer's pinned memory
            sink = BatchWriter(n, (h, w, 3)) if "sink" not in state else state["sink"]
            sink.rewind()
            state["sink"] = sink
            Embedder(BatchReader(clip), enc, sink, batch_frames=32).start()
            reader = BatchReader(sink.array)
        else:               # the reference's read() / write() per frame on pageable arrays
            if "sink" not in state:
                state["sink"] = Sink()
            sink = state["sink"]
            sink.k = 0
            Embedder(ArrayReader(list(clip)), enc, sink, batch_frames=32).start()
            reader = ArrayReader(list(sink.array))
        ex = Extractor(reader, DwtDctSvdDecoder(), DeShuffler(key=KEY).set_shape((PAYLOAD_LEN,)), batch_frames=32)
        ex.start()
        return ex.patterns
    state = {}
    fps, _, patterns = _wall_fps(flow, n, world, dev, 2)
    ok = float(np.mean([np.array_equal(p, PAYLOAD) for p in patterns]))
    how = ("BatchReader / BatchWriter over pinned arrays (optional batch protocol: whole batches move from / into the reader's and "
           "writer's memory)" if pinned_io else
           "pageable rgb24 numpy frames through per-frame read() / write() (uploads staged through pinned buffers by copy threads, "
           "three batches in flight: gather, GPU, write; the sink copies every frame into a preallocated array, as a pipe would)")
    return {"value": fps, "unit": "frames/s", "frames_per_gpu": n, "steps": 2, "h2d_bytes_per_step": 2 * n * h * w * 3,
            "d2h_bytes_per_step": n * h * w * 3 + n * PAYLOAD_LEN, "frames_exact": ok,
            "path": "offmark_b200 Embedder(batch_frames=32).start() + Extractor(batch_frames=32).start(), fused rgb24 kernels, " + how}


# --------------------------------------------------------------------------------- BASELINE configs[2]: 4K
def leg_4k(ops, dev, rank, world, peak, n=256, steps=10):
    h, w = 2160, 3840
    b = Batch(ops, dev, rank, world, n, h, w)
    marks = []
    for _ in range(3):
        b.step()
    b.finish()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    a.record()
    for _ in range(steps):
        vote = b.step(marks)
    b.finish()
    z.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(z), dev, world) / steps
    k_embed = float(np.mean([e[0].elapsed_time(e[1]) for e in marks]))
    k_extract = float(np.mean([e[1].elapsed_time(e[2]) for e in marks]))
    res = vote.result()
    seg_ok = sum(int(res[rank * b.n_seg_local + s][0] is not None and np.array_equal(res[rank * b.n_seg_local + s][0], b.payloads[s]))
                 for s in range(b.n_seg_local))
    embed_gbs = 2.0 * w * h * n / (k_embed * 1e-3) / 1e9
    return {"metric": "frames_per_sec_4k_embed_extract", "value": n * world / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "frames_per_gpu": n, "steps": steps, "ms_per_step": ms, "scaling": "weak",
            "config": "BASELINE configs[2]: 3840x2160 I420 frames, contiguous frame shards per rank, embed + extract + vote on Y",
            "roofline": {"bound": "hbm", "kernel": "dwtsvd_embed_tma_kernel (whole 30 KB strips, eight consumer warps)", "achieved": embed_gbs, "peak": peak,
                         "unit": "GB/s", "frac": embed_gbs / peak, "algorithmic_bytes_per_launch": 2 * w * h * n},
            "kernels": {"embed_ms": k_embed, "extract_ms": k_extract, "extract_GBs": w * h * n / (k_extract * 1e-3) / 1e9,
                        "step_GBs": 3.0 * w * h * n / (ms * 1e-3) / 1e9, "step_frac_of_peak": 3.0 * w * h * n / (ms * 1e-3) / 1e9 / peak},
            "segments_exact": seg_ok / b.n_seg_local}


# --------------------------------------------------------------------------------- the 8x8-DCT plugin pair
def leg_dct8(ops, batch, dev, peak, pool, n=512):
    """DctEncoder / DctDecoder (embed/dct_encoder.py:18-102, extract/dct_decoder.py:10-89) on 4:4:4 planar uint8: masks from
    the Y plane, mark in a full-resolution U plane.  Kernel times, roofline, a parity sample against the oracle and the
    reference-shaped CPU cost beside it."""
    from oracle import dct8 as o_dct, payload as o_pay
    h, w = batch.h, batch.w
    n = min(n, batch.n)
    yp = batch.src[:n]
    gen = torch.Generator(device=dev).manual_seed(SEED + 77)
    up = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    xx = torch.arange(w, device=dev, dtype=torch.float32)[None, None, :]
    yy = torch.arange(h, device=dev, dtype=torch.float32)[None, :, None]
    for f0 in range(0, n, 64):
        m = min(64, n - f0)
        u = 128 + 30 * torch.sin(xx / 53.0 + f0) * torch.cos(yy / 37.0) + 5.0 * torch.randn((m, h, w), device=dev, generator=gen)
        up[f0:f0 + m] = u.round().clamp(16, 240).to(torch.uint8)
    src_u = up.clone()
    wm_row = batch.rows[0]
    packed, ln = ops.pack_bits(wm_row, device=dev)
    masks = ops.dct8_masks(yp)
    ms_m = timed_ms(lambda: ops.dct8_masks(yp))
    ms_e = timed_ms(lambda: (up.copy_(src_u), ops.dct8_embed_(up, masks, packed, ln, alpha=20))) - timed_ms(lambda: up.copy_(src_u))
    up.copy_(src_u)
    ops.dct8_embed_(up, masks, packed, ln, alpha=20)
    ms_x = timed_ms(lambda: ops.dct8_extract(up, masks, alpha=20, payload_len=PAYLOAD_LEN))
    # the pair as the reference calls it: masks + quantiser per call (b200wm_dct8_encode / _decode, no caller-visible mask arrays)
    t_copy = timed_ms(lambda: up.copy_(src_u))
    ms_enc = timed_ms(lambda: (up.copy_(src_u), ops.dct8_encode_(yp, up, packed, ln, alpha=20))) - t_copy
    ms_dec = timed_ms(lambda: ops.dct8_decode(yp, up, alpha=20, payload_len=PAYLOAD_LEN))
    raw, counts = ops.dct8_decode(yp, up, alpha=20, payload_len=PAYLOAD_LEN)
    patterns, _ = batch.deg.degenerate_counts(counts, batch.block_num)
    payload = batch.payloads[0]
    frames_exact = float((patterns.cpu().numpy() == payload).all(axis=1).mean())
    # parity sample: a 256-row band of frame 0, run as its own plane on both sides (the luminance mask depends on the
    # frame-global mean, dct_encoder.py:54-56), through the reference-shaped oracle
    band = 256
    y0, u0 = yp[0, :band].cpu().numpy(), src_u[0, :band].cpu().numpy()
    yuv = np.zeros((band, w, 3), dtype=np.float32)
    yuv[:, :, 0], yuv[:, :, 1] = y0, u0
    want = np.around(np.clip(o_dct.encode(yuv.copy(), wm_row[None, :])[:, :, 1], 0, 255)).astype(np.uint8)
    t_y, t_u = torch.from_numpy(y0).to(dev), torch.from_numpy(u0.copy()).to(dev)
    ops.dct8_encode_(t_y, t_u, packed, ln, alpha=20)
    got = t_u.cpu().numpy()
    d = np.abs(got.astype(np.int16) - want)
    rawb, _ = ops.dct8_decode(t_y, t_u, alpha=20)
    bits = ops.unpack_bits(rawb, band * w // 64)[0]
    yuv[:, :, 1] = got
    bits_ref = o_dct.decode(yuv)[0].astype(np.uint8)
    blocks_differ = int((d.reshape(band // 8, 8, w // 8, 8).max(axis=(1, 3)) > 1).sum())
    cpu = None
    if pool is not None:
        cores = pool.cores
        ys = batch.src[:cores, :64].cpu().numpy()
        us = src_u[:cores, :64].cpu().numpy()
        t0 = time.perf_counter()
        pool.map(_cpu_worker_dct8, [(ys[i:i + 1], us[i:i + 1], wm_row) for i in range(cores)])
        dt = time.perf_counter() - t0
        cpu = {"value": cores * 64 / h / dt, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{cores} bands of 64x{w}, DctEncoder.encode + DctDecoder.decode + vote as the reference runs them (per-block cv2.dct loops), {dt:.1f} s"}
    step_ms = ms_enc + ms_dec                    # encode + decode, each with its own masks, as the reference runs them
    total_bytes = (3 + 2) * w * h * n            # encode: Y + U in + U out; decode: Y + U
    return {"metric": "frames_per_sec_1080p_dct8_embed_extract", "value": n / (step_ms * 1e-3), "unit": "frames/s", "frames": n,
            "layout": "planar uint8 4:4:4 (masks from Y, mark in U)",
            "calls": "b200wm_dct8_encode + b200wm_dct8_decode (masks + quantiser per call, no caller-visible mask arrays); the "
                     "kernels behind them are timed beside it",
            "roofline": {"bound": "hbm", "kernel": "dct8_masks_kernel", "achieved": w * h * n / (ms_m * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": w * h * n / (ms_m * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": w * h * n,
                         "note": "this kernel is FP32-pipe bound, not HBM bound: 640 FP32 operations per 64-byte block (full 2-D DCT "
                                 "butterflies + magnitude sums) put its ceiling at 0.57 of the HBM peak at 100 % FP32 issue (DESIGN.md 4.4)"},
            "kernels": {"encode_ms": ms_enc, "encode_frac": 3 * w * h * n / (ms_enc * 1e-3) / 1e9 / peak,
                        "decode_ms": ms_dec, "decode_frac": 2 * w * h * n / (ms_dec * 1e-3) / 1e9 / peak,
                        "masks_ms": ms_m, "masks_frac": w * h * n / (ms_m * 1e-3) / 1e9 / peak,
                        "embed_ms": ms_e, "embed_frac": 2 * w * h * n / (ms_e * 1e-3) / 1e9 / peak,
                        "extract_ms": ms_x, "extract_frac": w * h * n / (ms_x * 1e-3) / 1e9 / peak,
                        "pair_GBs": total_bytes / (step_ms * 1e-3) / 1e9, "pair_frac_of_peak": total_bytes / (step_ms * 1e-3) / 1e9 / peak},
            "frames_exact": frames_exact,
            "parity": {"sample": f"{band}x{w} band of frame 0 as its own plane, against oracle/dct8.py", "embed_max_abs_lsb": int(d.max()),
                       "embed_frac_pixels_diff": float((d > 0).mean()), "blocks_more_than_1_lsb_off": blocks_differ,
                       "raw_bit_agreement": float((bits == bits_ref).mean()),
                       "voted_payload_identical": bool(np.array_equal(o_pay.degenerate(bits.reshape(1, -1).astype(np.float64), PAYLOAD_LEN, KEY),
                                                                     o_pay.degenerate(bits_ref.reshape(1, -1).astype(np.float64), PAYLOAD_LEN, KEY)))},
            "cpu_baseline": cpu}


# --------------------------------------------------------------------------------- BASELINE configs[4]: attacks
def leg_attacks(args, ops, batch, dev, pool):
    """Detection after distortions over ``--attack-frames`` marked 1080p frames (10,000 by default): every frame is
    generated, marked, attacked and read back on the GPU; on ``--attack-oracle-frames`` (64) of the very same attacked
    frames per attack the reference extractor (oracle) reads them too."""
    h, w = batch.h, batch.w
    n = args.attack_frames
    block_num = batch.block_num
    wm_row = batch.rows[0]
    payload = batch.payloads[0]
    packed, ln = ops.pack_bits(wm_row, device=dev)
    truth = torch.from_numpy(wm_row.astype(np.uint8)).to(dev)
    t0 = time.perf_counter()
    marked = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    for f0 in range(0, n, 1000):
        m = min(1000, n - f0)
        marked[f0:f0 + m] = y_planes(generate_i420(m, dev, 1000 + f0 // 1000), h, w)
    ops.dwtsvd_embed_(marked, packed, ln)
    work = torch.empty_like(marked)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0

    def noise(sigma):
        g = torch.Generator(device=dev).manual_seed(int(sigma * 1000))
        for f0 in range(0, n, 100):
            m = min(100, n - f0)
            ops.attack_add_noise_(work[f0:f0 + m], sigma * torch.randn((m, h, w), device=dev, generator=g))

    def resize():
        for f0 in range(0, n, 500):
            ops.attack_resize_roundtrip_(work[f0:f0 + 500])

    attacks = [("none", lambda: None), ("jpeg_like_q95", lambda: ops.attack_jpeg_requant_(work, 95)),
               ("jpeg_like_q85", lambda: ops.attack_jpeg_requant_(work, 85)), ("jpeg_like_q75", lambda: ops.attack_jpeg_requant_(work, 75)),
               ("gaussian_sigma_1", lambda: noise(1.0)), ("gaussian_sigma_2", lambda: noise(2.0)), ("gaussian_sigma_4", lambda: noise(4.0)),
               ("resize_1080p_720p_1080p", resize)]
    k = min(args.attack_oracle_frames, n)
    idx = np.linspace(0, n - 1, k).astype(int)
    sel = torch.from_numpy(idx).to(dev)
    pay_t = torch.from_numpy(payload.astype(np.uint8)).to(dev)
    raw = torch.empty((n, batch.words), dtype=torch.int32, device=dev)
    counts = torch.empty((n, PAYLOAD_LEN), dtype=torch.int32, device=dev)
    table = []
    for name, attack in attacks:
        work.copy_(marked)
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        attack()
        z.record()
        ops.dwtsvd_extract(work, payload_len=PAYLOAD_LEN, raw_bits=raw, pos_counts=counts)
        patterns, _ = batch.deg.degenerate_counts(counts, block_num)
        torch.cuda.synchronize()
        # raw BER vs the embedded bits over ALL frames, on the GPU (plumbing: bit unpacking with torch ops, in chunks)
        wrong = 0
        for f0 in range(0, n, 1000):
            r = raw[f0:f0 + 1000].view(torch.uint8)
            b = ((r[:, :, None] >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).reshape(r.shape[0], -1)[:, :block_num]
            wrong += int((b != truth[None, :]).sum())
        row = {"attack": name, "frames": n, "attack_ms": a.elapsed_time(z), "raw_ber_b200": wrong / (n * block_num),
               "payload_exact_b200": float((patterns == pay_t).all(dim=1).float().mean())}
        if pool is not None and k:
            host = work[sel].cpu().numpy()
            bits = ops.unpack_bits(raw[sel], block_num)
            pats = patterns[sel].cpu().numpy()
            res = pool.map(_attack_worker, [(host[j], bits[j], pats[j], wm_row) for j in range(k)])
            nbits = sum(r[4] for r in res)
            row.update({"oracle_frames": k, "raw_ber_reference": float(np.mean([r[0] for r in res])),
                        "raw_ber_b200_same_frames": float(np.mean([(bits[j] != wm_row).mean() for j in range(k)])),
                        "raw_bit_agreement": 1.0 - sum(r[1] for r in res) / nbits, "raw_bit_mismatches": sum(r[1] for r in res),
                        "raw_bit_mismatches_off_boundary": sum(r[2] for r in res), "knife_edge_blocks": sum(r[3] for r in res),
                        "voted_payload_identical": float(np.mean([r[5] for r in res]))})
        table.append(row)
    del marked, work, raw
    return {"config": "BASELINE configs[4]: detection after distortions", "frames": n, "generate_and_mark_s": t_gen, "table": table}


if __name__ == "__main__":
    main()
