"""Segment / copy fingerprint layer over the batched kernels.

The reference builds an HLS ladder in which every 2-second segment exists in N differently marked
copies; a player's copy sequence is its fingerprint (tests/mark_video_to_hls.py, tests/detect_watermarks.py).
This module keeps the reference's payload schemes, decision rules and JSON shapes and replaces its
per-copy, per-frame Python loops by one ``b200wm_dwtsvd_embed_copies`` launch per batch of frames (one read,
N marked copies) and one extract + vote launch per batch on the way back.  ffmpeg segmentation, encoding
and playlist writing stay outside.
"""
import numpy as np
import torch

from b200wm import ops
from b200wm.vote import SegmentVote
from .generator.shuffler import Shuffler
from .degenerator.de_shuffler import DeShuffler


def generate_payload_for_segment(segment_number, copy_index=None):
    """``copy_index is None``: 8-bit segment number (tests/segment_mark_detect_hls.py:42-55).
    Otherwise 4-bit segment number || 4-bit copy index, both modulo 16 (tests/mark_video_to_hls.py:27-43)."""
    if copy_index is None:
        binary = format(int(segment_number) % 256, '08b')
    else:
        binary = format(int(segment_number) % 16, '04b') + format(int(copy_index) % 16, '04b')
    return np.array([int(bit) for bit in binary])


def decode_watermark_pattern(pattern):
    """(segment_number, copy_index) of an 8-bit pattern, (None, None) otherwise (tests/detect_watermarks.py:145-172)."""
    if pattern is None:
        return None, None
    binary_str = ''.join(map(str, [int(b) for b in pattern]))
    if len(binary_str) >= 8:
        return int(binary_str[:4], 2), int(binary_str[4:8], 2)
    return None, None


def mark_segment_copies(planes, frame_segment_number, n_copies, key=0, scale=15.0, out=None):
    """N marked copies of every frame from ONE read of the source (mark_video_to_hls.py:330-354).

    ``planes`` uint8 CUDA ``[F, H, W]`` (luma or a chroma plane); ``frame_segment_number`` the segment number of
    every frame (sequence of F ints).  Returns ``(copies [n_copies, F, H, W], segment_payloads, segment_copies)``
    where the two dicts have the shapes the reference writes to ``segment_payloads.json`` / ``segment_copies.json``
    (:339-354; the ``file`` entries follow its naming ``marked_seg{segment}_copy{copy}.mp4``)."""
    segs = [int(s) for s in frame_segment_number]
    if len(segs) != planes.shape[0]:
        raise ValueError("one segment number per frame")
    order = sorted(set(segs))
    h, w = planes.shape[-2], planes.shape[-1]
    capacity = (1, h * w // 64)
    rows, segment_payloads, segment_copies = [], {}, {}
    for s in order:
        segment_copies[str(s)] = []
        for c in range(n_copies):
            payload = generate_payload_for_segment(s, c)
            rows.append(Shuffler(key=key).generate_wm(payload.copy(), capacity)[0])
            segment_payloads[f"{s}_{c}"] = payload.tolist()
            segment_copies[str(s)].append({"file": f"marked_seg{s}_copy{c}.mp4", "payload": payload.tolist(), "copy_index": c})
    packed, n = ops.pack_bits(np.stack(rows), device=planes.device)
    index = {s: i for i, s in enumerate(order)}
    table = torch.tensor([[index[s] * n_copies + c for c in range(n_copies)] for s in segs], dtype=torch.int32, device=planes.device)
    copies = ops.dwtsvd_embed_copies(planes, packed, n, n_copies, scale=scale, copy_wm_row=table, out=out)
    return copies, segment_payloads, segment_copies


def detect_segment_copies(planes, frame_segment_number, segment_payloads=None, max_copies=16, key=0, scale=15.0, group=None):
    """Which copy of every segment is this?  Per-frame extract + vote on the GPU, then the decision rules of
    tests/detect_watermarks.py:325-365: with a payload map, the copy whose expected payload equals the segment's
    most common pattern (highest frequency wins); without one, the pattern is decoded and accepted when its
    segment bits match ``segment_number % 16``.  Returns the list the reference dumps to ``detection_results.json``
    (:374-383) - ``segment`` holds the segment number instead of a file name."""
    segs = [int(s) for s in frame_segment_number]
    order = sorted(set(segs))
    index = {s: i for i, s in enumerate(order)}
    h, w = planes.shape[-2], planes.shape[-1]
    deg = DeShuffler(key=key).set_shape((8,))
    _, counts = ops.dwtsvd_extract(planes, scale=scale, payload_len=8)
    _, packed = deg.degenerate_counts(counts, h * w // 64)
    frame_seg = torch.tensor([index[s] for s in segs], dtype=torch.int32, device=planes.device)
    vote = SegmentVote(len(order), 8, planes.device).add(packed, frame_segment=frame_seg).combine(group)
    results = []
    for s, (pattern, frequency, _, frames) in zip(order, vote.result()):
        detected_copy, best = None, 0
        if pattern is not None and segment_payloads:
            for copy_index in range(max_copies):
                expected = segment_payloads.get(f"{s}_{copy_index}")
                if expected is None:
                    continue
                if np.array_equal(pattern, np.array(expected)) and frequency > best:
                    best, detected_copy = frequency, copy_index
        elif pattern is not None:
            seg_bits, copy_bits = decode_watermark_pattern(pattern)
            if seg_bits is not None and seg_bits == s % 16:
                detected_copy, best = copy_bits, frequency
        results.append({"segment": s, "segment_number": s, "detected_copy_index": detected_copy,
                        "match_frequency": best, "success": detected_copy is not None})
    return results


def copy_fingerprint(results):
    """Copy sequence in segment order and, when every segment was identified, the fingerprint string
    (tests/detect_watermarks.py:404-425)."""
    ordered = sorted(results, key=lambda r: r["segment_number"] if r["segment_number"] is not None else float("inf"))
    sequence = [r["detected_copy_index"] for r in ordered]
    fingerprint = ''.join(str(c) for c in sequence) if all(c is not None for c in sequence) else None
    return sequence, fingerprint
