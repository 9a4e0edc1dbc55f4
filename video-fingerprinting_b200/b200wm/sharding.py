"""How frames and HLS segments are dealt to the GPUs of one box (one process per GPU).

Embed and raw-bit extract are independent per frame, the per-frame vote is frame-local and
only the per-segment pattern vote crosses frames, so the data path needs no
collective: every rank works on its own shard and only the vote counters are combined
(``b200wm.vote.SegmentVote.combine``)."""


def frame_shard(n_frames, rank, world):
    """Contiguous chunk ``[start, stop)`` of a batch of independent frames."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def segment_shard(n_segments, rank, world):
    """Whole segments dealt round-robin: segment ``s`` goes to rank ``s % world``."""
    return list(range(rank, n_segments, world))
