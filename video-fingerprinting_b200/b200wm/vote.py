"""Cross-frame / cross-GPU vote (host logic over the device counters).

Reference semantics (tests/segment_mark_detect_hls.py:144-155, duplicated at
tests/mark_video_to_hls.py:273-282 and tests/detect_watermarks.py:126-137): the per-frame
patterns of a segment are joined to strings, ``Counter(...).most_common(1)`` picks the mode
(ties go to the pattern that appeared first) and ``frequency = count / frames``.

Frames and segments shard across GPUs (one process per GPU).  Each rank accumulates, with
``b200wm_pattern_hist``, a histogram of patterns per segment, the earliest global frame index
of each pattern, the per-bit vote counters and the frame counts; ``combine`` merges the ranks
with one all-gather of the flat state (a few tens of KB over NCCL/NVLink; SUM of the counters and
MIN of the first-seen indices are then taken locally), after which every rank can reproduce
``most_common(1)`` exactly.
"""
from collections import Counter

import numpy as np
import torch
import torch.distributed as dist

from . import ops

MAX_HIST_BITS = 16


class SegmentVote:
    def __init__(self, n_segments, payload_len, device):
        if payload_len > MAX_HIST_BITS:
            raise ValueError(f"pattern histogram supports payload_len <= {MAX_HIST_BITS}; "
                             "use gathered_pattern_vote for longer payloads")
        self.n_segments, self.payload_len = int(n_segments), int(payload_len)
        dev = torch.device(device)
        bins = 1 << self.payload_len
        # one flat int32 buffer [hist | bit_votes | seg_frames | first_seen] so that ONE collective moves it all
        a, b = self.n_segments * bins, self.n_segments * (bins + self.payload_len)
        c = b + self.n_segments
        self._flat = torch.zeros(c + self.n_segments * bins, dtype=torch.int32, device=dev)
        self._n_sum = c
        self.hist = self._flat[:a].view(self.n_segments, bins)
        self.bit_votes = self._flat[a:b].view(self.n_segments, self.payload_len)
        self.seg_frames = self._flat[b:c].view(self.n_segments)
        self.first_seen = self._flat[c:].view(self.n_segments, bins)
        self.first_seen.fill_(ops.INT32_MAX)

    def _state(self):
        return {"hist": self.hist, "first_seen": self.first_seen, "bit_votes": self.bit_votes,
                "seg_frames": self.seg_frames}

    def add(self, packed, frame_segment=None, frame_order=None, order_offset=0):
        """Accumulate per-frame packed patterns (int64 ``[N]`` from ``ops.vote_finish``) on the GPU."""
        ops.pattern_hist(packed, self.payload_len, self.n_segments, frame_segment, frame_order, order_offset,
                         state=self._state())
        return self

    def combine(self, group=None):
        """Merge the ranks' counters: one all-gather of the flat state (tens of KB per rank, latency-bound
        over NVLink), then SUM of the counters and MIN of the first-seen indices locally.  No-op
        without an initialised process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            world = dist.get_world_size(group)
            parts = [torch.empty_like(self._flat) for _ in range(world)]
            dist.all_gather(parts, self._flat, group=group)
            stacked = torch.stack(parts)
            self._flat[:self._n_sum] = stacked[:, :self._n_sum].sum(dim=0, dtype=torch.int32)
            self._flat[self._n_sum:] = stacked[:, self._n_sum:].amin(dim=0)
        return self

    def result(self):
        """Per segment: (pattern uint8 [L] or None, frequency or None, bit_votes int [L], frames)."""
        hist = self.hist.cpu().numpy()
        first = self.first_seen.cpu().numpy()
        votes = self.bit_votes.cpu().numpy()
        frames = self.seg_frames.cpu().numpy()
        out = []
        for s in range(self.n_segments):
            if frames[s] == 0:
                out.append((None, None, votes[s].copy(), 0))       # "No patterns collected" (:140-142)
                continue
            top = hist[s].max()
            tied = np.flatnonzero(hist[s] == top)
            best = tied[np.argmin(first[s][tied])]                 # Counter: first seen wins ties
            pattern = np.array([(int(best) >> (self.payload_len - 1 - j)) & 1 for j in range(self.payload_len)],
                               dtype=np.uint8)
            out.append((pattern, float(top) / float(frames[s]), votes[s].copy(), int(frames[s])))
        return out


def gathered_pattern_vote(patterns, frame_order=None, group=None):
    """Reference-exact vote for any payload length: all-gather the per-frame patterns
    (uint8 ``[N, L]`` tensor, a few bytes per frame) with their global frame order, then take
    the ``Counter`` mode on the host exactly as the reference does."""
    n, length = patterns.shape
    if frame_order is None:
        frame_order = torch.arange(n, dtype=torch.int64, device=patterns.device)
    order = frame_order.to(torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        sizes = [torch.zeros(1, dtype=torch.int64, device=patterns.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=patterns.device), group=group)
        cap = int(max(int(s.item()) for s in sizes))
        pad_p = torch.zeros((cap, length), dtype=torch.uint8, device=patterns.device)
        pad_o = torch.full((cap,), -1, dtype=torch.int64, device=patterns.device)
        pad_p[:n], pad_o[:n] = patterns, order
        all_p = [torch.empty_like(pad_p) for _ in range(world)]
        all_o = [torch.empty_like(pad_o) for _ in range(world)]
        dist.all_gather(all_p, pad_p, group=group)
        dist.all_gather(all_o, pad_o, group=group)
        p = torch.cat(all_p).cpu().numpy()
        o = torch.cat(all_o).cpu().numpy()
        keep = o >= 0
        p, o = p[keep], o[keep]
    else:
        p, o = patterns.cpu().numpy(), order.cpu().numpy()
    if len(p) == 0:
        return None, None
    p = p[np.argsort(o, kind="stable")]
    strings = [''.join(map(str, row)) for row in p]
    best, count = Counter(strings).most_common(1)[0]
    return np.array([int(ch) for ch in best], dtype=np.uint8), count / len(strings)
