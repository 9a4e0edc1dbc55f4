"""Cross-frame / cross-GPU vote (host logic over the device counters).

Reference semantics (tests/segment_mark_detect_hls.py:144-155, duplicated at
tests/mark_video_to_hls.py:273-282 and tests/detect_watermarks.py:126-137): the per-frame
patterns of a segment are joined to strings, ``Counter(...).most_common(1)`` picks the mode
(ties go to the pattern that appeared first) and ``frequency = count / frames``.

Frames and segments shard across GPUs (one process per GPU).  Each rank accumulates, with
``b200wm_pattern_hist``, a histogram of patterns per segment, the earliest global frame index
of each pattern, the per-bit vote counters and the frame counts; ``combine`` merges the ranks
with ONE all-gather over NCCL/NVLink - of each rank's own block when whole segments are dealt to
ranks (nothing to reduce), else of the full state (SUM of the counters and MIN of the first-seen
indices are then taken locally) - after which every rank can reproduce ``most_common(1)`` exactly.
"""
from collections import Counter

import numpy as np
import torch
import torch.distributed as dist

from . import ops

MAX_HIST_BITS = 16


class SegmentVote:
    """Per-segment vote state on one rank.

    ``owned=(first, count)`` declares that this rank only ever adds frames of the ``count`` consecutive
    segments starting at ``first`` (whole HLS segments dealt to ranks, BASELINE config 4) and that every rank
    owns the same number of segments in rank order.  The state is then laid out rank-major, ``add`` takes
    GLOBAL segment numbers as before, and ``combine`` is ONE in-place all-gather of this rank's block with
    nothing to reduce afterwards.  Without ``owned`` a segment may be split across ranks and ``combine``
    all-gathers the full state and reduces it locally (SUM of counters, MIN of first-seen indices)."""

    def __init__(self, n_segments, payload_len, device, owned=None, symmetric=False, group=None):
        if payload_len > MAX_HIST_BITS:
            raise ValueError(f"pattern histogram supports payload_len <= {MAX_HIST_BITS}; "
                             "use gathered_pattern_vote for longer payloads")
        self.n_segments, self.payload_len = int(n_segments), int(payload_len)
        dev = torch.device(device)
        bins = 1 << self.payload_len
        self.owned = None
        if owned is not None:
            first, count = int(owned[0]), int(owned[1])
            if count <= 0 or self.n_segments % count or first % count or first + count > self.n_segments:
                raise ValueError("owned=(first, count) needs equal blocks of consecutive segments in rank order")
            self.owned = (first, count)
        blocks = self.n_segments // self.owned[1] if self.owned else 1
        per = self.owned[1] if self.owned else self.n_segments           # segments per block
        # per block one flat int32 run [hist | bit_votes | seg_frames | first_seen]: ONE collective moves a block
        a, b = per * bins, per * (bins + self.payload_len)
        c = b + per
        self._block_len = c + per * bins
        self._n_sum = c
        self._work = None
        self._gathered = None
        self._symm = None
        if symmetric:
            # ``symmetric=True`` (owned mode, inside a process group): the state lives in peer-mapped memory and ``add``
            # runs the histogram kernel FUSED with the NVLink exchange of this rank's block (b200wm_pattern_hist_publish):
            # no collective call at all, ``combine`` has nothing left to do.  Collective constructor (rendezvous).
            # The exchange is one-sided: a peer's next ``add`` to the SAME state object overwrites its slot in this rank's
            # buffer, so a state must have been read (``result()``) on every rank before any rank re-uses it - use two
            # state objects alternately (as bench.py does) when batches are pipelined.
            if not self.owned or not (dist.is_available() and dist.is_initialized()):
                raise ValueError("symmetric=True needs owned=(first, count) and an initialised process group")
            import torch.distributed._symmetric_memory as symm_mem
            self._pad_len = (self._block_len + 3) // 4 * 4                 # 16-byte blocks for the 128-bit peer stores
            world = dist.get_world_size(group)
            if blocks != world:
                raise ValueError("owned blocks must follow the rank order of the group")
            buf = symm_mem.empty(world * self._pad_len + max(4, (world + 3) // 4 * 4), dtype=torch.int32, device=dev)
            buf.zero_()
            handle = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            self._symm = {"buf": buf, "handle": handle, "peers": int(handle.buffer_ptrs_dev), "world": world,
                          "rank": dist.get_rank(group), "epoch": 0,
                          "ticket": torch.zeros(2, dtype=torch.int32, device=dev), "status": torch.zeros(1, dtype=torch.int32, device=dev)}
            torch.cuda.synchronize(dev)
            dist.barrier(group)                                            # every peer's flags are zero before anyone publishes
            self._flat = buf[:world * self._pad_len]
            blk = self._flat.view(world, self._pad_len)[:, :self._block_len]
        else:
            self._flat = torch.zeros(blocks * self._block_len, dtype=torch.int32, device=dev)
            blk = self._flat.view(blocks, self._block_len)
        self.hist = blk[:, :a].view(blocks, per, bins)
        self.bit_votes = blk[:, a:b].view(blocks, per, self.payload_len)
        self.seg_frames = blk[:, b:c].view(blocks, per)
        self.first_seen = blk[:, c:].view(blocks, per, bins)
        self.first_seen.fill_(ops.INT32_MAX)
        if not self.owned:                # one block: the plain [n_segments, ...] views of the general mode
            self.hist, self.bit_votes = self.hist[0], self.bit_votes[0]
            self.seg_frames, self.first_seen = self.seg_frames[0], self.first_seen[0]

    def reset(self):
        """Back to "nothing seen" so that ONE state object serves batch after batch: a single kernel launch on the
        GPU (``b200wm_vote_state_reset``), no allocation.  Waits for a pending asynchronous ``combine`` first."""
        if self._symm is None:
            self.wait()                     # (symmetric mode: one-sided exchange, only this rank's own block is reset)
        if self._symm is not None:
            k = self.owned[0] // self.owned[1]
            ops.vote_state_reset(self._flat[k * self._pad_len:k * self._pad_len + self._block_len], self._n_sum)
        elif self._flat.is_cuda:
            if self.owned:          # per block [counters | first_seen]: reset block by block views of the same run
                blocks = self.n_segments // self.owned[1]
                if blocks == 1:
                    ops.vote_state_reset(self._flat, self._n_sum)
                else:               # only this rank's block is ever accumulated into; the others are overwritten by the gather
                    k = self.owned[0] // self.owned[1]
                    ops.vote_state_reset(self._flat[k * self._block_len:(k + 1) * self._block_len], self._n_sum)
            else:
                ops.vote_state_reset(self._flat, self._n_sum)
        else:
            blocks = self._flat.numel() // self._block_len
            blk = self._flat.view(blocks, self._block_len)
            blk[:, :self._n_sum] = 0
            blk[:, self._n_sum:] = ops.INT32_MAX
        return self

    def _mine(self):
        """State views that ``b200wm_pattern_hist`` accumulates into and the segment offset it subtracts."""
        if not self.owned:
            return {"hist": self.hist, "first_seen": self.first_seen, "bit_votes": self.bit_votes,
                    "seg_frames": self.seg_frames}, 0
        k = self.owned[0] // self.owned[1]
        return {"hist": self.hist[k], "first_seen": self.first_seen[k], "bit_votes": self.bit_votes[k],
                "seg_frames": self.seg_frames[k]}, self.owned[0]

    def _state(self):
        return self._mine()[0]

    def add(self, packed, frame_segment=None, frame_order=None, order_offset=0):
        """Accumulate per-frame packed patterns (int64 ``[N]`` from ``ops.vote_finish``) on the GPU.
        ``frame_segment`` holds global segment numbers (owned mode: all inside this rank's block)."""
        state, first = self._mine()
        n_seg = self.owned[1] if self.owned else self.n_segments
        if frame_segment is not None and frame_segment.dtype != torch.int32:
            frame_segment = frame_segment.to(torch.int32)
        if frame_order is not None and frame_order.dtype != torch.int32:
            frame_order = frame_order.to(torch.int32)
        if first and frame_segment is not None:
            frame_segment = frame_segment - first
        if self._symm is not None:
            s = self._symm
            # the fused kernel takes at most 128 - 4 (world - 1) CTAs of 256 frames (its helper CTAs spin, so the whole grid
            # must be resident): longer batches accumulate their head with the plain histogram kernel
            limit = (128 - 4 * (s["world"] - 1)) * 256
            n = packed.numel()
            if n > limit:
                head = n - limit
                ops.pattern_hist(packed[:head], self.payload_len, n_seg, None if frame_segment is None else frame_segment[:head].contiguous(),
                                 None if frame_order is None else frame_order[:head].contiguous(), order_offset, state=state)
                packed = packed[head:]
                frame_segment = None if frame_segment is None else frame_segment[head:].contiguous()
                frame_order = None if frame_order is None else frame_order[head:].contiguous()
                order_offset += head
            s["epoch"] += 1
            s["pending"] = True
            ops.pattern_hist_publish(packed, self.payload_len, n_seg, state, frame_segment, frame_order, order_offset, s["peers"],
                                     self._pad_len, s["world"], s["rank"], s["epoch"], s["ticket"], s["status"])
            return self
        ops.pattern_hist(packed, self.payload_len, n_seg, frame_segment, frame_order, order_offset, state=state)
        return self

    def combine(self, group=None, async_op=False):
        """Merge the ranks' counters (tens of KB per rank, latency-bound over NVLink).  Owned mode: one
        in-place all-gather of this rank's block.  General mode: one all-gather of the flat state, then SUM
        of the counters and MIN of the first-seen indices locally.  No-op without a process group.

        ``async_op=True`` enqueues the collective behind the work already on the current stream and returns
        at once WITHOUT making the current stream wait for it, so that the next batch's embed overlaps the
        exchange; ``wait()`` (called by ``result`` and ``reset``) joins it."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return self
        if self._symm is not None:          # exchanged by the kernel that accumulated it
            return self
        self.wait()
        world = dist.get_world_size(group)
        if self.owned:
            if world * self.owned[1] != self.n_segments or dist.get_rank(group) != self.owned[0] // self.owned[1]:
                raise ValueError("owned blocks must follow the rank order of the group")
            k = self.owned[0] // self.owned[1]
            mine = self._flat[k * self._block_len:(k + 1) * self._block_len]
            work = dist.all_gather_into_tensor(self._flat, mine, group=group, async_op=async_op)
        else:
            if self._gathered is None or self._gathered.shape[0] != world:
                self._gathered = torch.empty((world, self._flat.numel()), dtype=torch.int32, device=self._flat.device)
            work = dist.all_gather_into_tensor(self._gathered.view(-1), self._flat, group=group, async_op=async_op)
        if async_op:
            self._work = work
        else:
            self._finish()
        return self

    def _finish(self):
        if not self.owned and self._gathered is not None:
            self._flat[:self._n_sum] = self._gathered[:, :self._n_sum].sum(dim=0, dtype=torch.int32)
            self._flat[self._n_sum:] = self._gathered[:, self._n_sum:].amin(dim=0)

    def wait(self):
        """Join a pending asynchronous ``combine`` (the current stream waits for the collective) or, in symmetric
        mode, enqueue the wait for the peers' blocks of the last exchange."""
        if self._symm is not None and self._symm.get("pending"):
            s = self._symm
            ops.vote_exchange_wait(s["peers"], self._pad_len, s["world"], s["rank"], s["epoch"], s["status"])
            s["pending"] = False
        if self._work is not None:
            self._work.wait()
            self._work = None
            self._finish()
        return self

    def result(self):
        """Per segment: (pattern uint8 [L] or None, frequency or None, bit_votes int [L], frames)."""
        self.wait()
        if self._symm is not None and int(self._symm["status"].item()):
            raise RuntimeError("vote exchange timed out: a peer did not publish its block within two seconds")
        bins = 1 << self.payload_len
        hist = self.hist.cpu().numpy().reshape(self.n_segments, bins)
        first = self.first_seen.cpu().numpy().reshape(self.n_segments, bins)
        votes = self.bit_votes.cpu().numpy().reshape(self.n_segments, self.payload_len)
        frames = self.seg_frames.cpu().numpy().reshape(self.n_segments)
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
