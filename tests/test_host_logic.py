"""CPU: host-side logic of the drop-in (generators, sharding, vote combine incl. world_size 2 on gloo)."""
import os
import socket
from collections import Counter

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import payload as o_pay
from parity import PAYLOAD


def test_shuffler_matches_oracle():
    from offmark_b200.generator.shuffler import Shuffler
    for key, length, cap in ((0, 8, (1, 32400)), (5, 8, (1, 1200)), (0, 5, (1, 97)), (11, 16, (1, 129600))):
        p = np.random.RandomState(key + 1).randint(0, 2, length)
        a = Shuffler(key=key).generate_wm(p, cap)
        b = o_pay.generate_wm(p, cap, key)
        assert a.shape == b.shape == cap and a.dtype == b.dtype and np.array_equal(a, b)
    assert Shuffler.wm_type() == "bits"
    wm = Shuffler(key=0).generate_wm(PAYLOAD, (1, 64))
    assert np.array_equal(wm[0, :8], PAYLOAD[[6, 2, 1, 7, 3, 0, 5, 4]])


def test_grayscale_matches_oracle():
    from offmark_b200.generator.grayscale import GrayScale
    img = np.random.RandomState(2).randint(0, 256, (21, 21)).astype(np.uint8)
    a = GrayScale(key=0).generate_wm(img, (1, 32400))
    assert np.array_equal(a, o_pay.generate_wm_grayscale(img, (1, 32400), 0))
    assert GrayScale.wm_type() == "grayscale"
    with pytest.warns(UserWarning):
        GrayScale(key=0).generate_wm(img, (1, 100))


def test_sharding():
    from b200wm.sharding import frame_shard, segment_shard
    for n in (0, 1, 7, 3000, 1024):
        for world in (1, 2, 4, 8):
            spans = [frame_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert segment_shard(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((segment_shard(64, r, 8) for r in range(8)), [])) == list(range(64))


def _fill_vote_state(vote, patterns_int, segments, orders):
    """Host-side stand-in for b200wm_pattern_hist so the combine logic can run without a GPU."""
    L = vote.payload_len
    for p, s, o in zip(patterns_int, segments, orders):
        vote.hist[s, p] += 1
        vote.first_seen[s, p] = min(int(vote.first_seen[s, p]), int(o))
        vote.seg_frames[s] += 1
        for j in range(L):
            vote.bit_votes[s, j] += (p >> (L - 1 - j)) & 1


def _reference_vote(patterns_int, L):
    strings = [format(p, f"0{L}b") for p in patterns_int]
    best, count = Counter(strings).most_common(1)[0]
    return [int(c) for c in best], count / len(strings)


def test_segment_vote_result_matches_counter_incl_ties():
    from b200wm.vote import SegmentVote
    rng = np.random.RandomState(0)
    L, S = 8, 5
    vote = SegmentVote(S, L, "cpu")
    per_seg = {}
    for s in range(S - 1):
        pats = rng.choice([0x65, 0x9A, 0x01], size=60 if s else 4, p=[0.4, 0.4, 0.2]).tolist()
        if s == 0:
            pats = [0x9A, 0x65, 0x65, 0x9A]          # exact tie: first seen (0x9A) must win
        per_seg[s] = pats
        _fill_vote_state(vote, pats, [s] * len(pats), range(len(pats)))
    res = vote.result()
    for s, pats in per_seg.items():
        want_p, want_f = _reference_vote(pats, L)
        assert res[s][0].tolist() == want_p and res[s][1] == want_f and res[s][3] == len(pats)
        assert res[s][2].tolist() == [sum((p >> (L - 1 - j)) & 1 for p in pats) for j in range(L)]
    assert res[S - 1][0] is None and res[S - 1][1] is None and res[S - 1][3] == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _vote_worker(rank, world, port, L, all_patterns, out_q):
    import sys
    from conftest import ROOT, PKG      # noqa: F401  (sets sys.path in the child)
    from b200wm.vote import SegmentVote, gathered_pattern_vote
    from b200wm.sharding import frame_shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = len(all_patterns)
    a, b = frame_shard(n, rank, world)
    mine = all_patterns[a:b]
    vote = SegmentVote(1, L, "cpu")
    _fill_vote_state(vote, mine, [0] * len(mine), range(a, b))
    vote.combine()
    first_round = vote.result()[0]
    # the same state object serves the next batch: reset, refill, asynchronous combine joined by result()
    vote.reset()
    assert vote.result()[0][0] is None
    _fill_vote_state(vote, mine, [0] * len(mine), range(a, b))
    vote.combine(async_op=True)
    pattern, freq, bit_votes, frames = vote.result()[0]
    assert pattern.tolist() == first_round[0].tolist() and freq == first_round[1] and frames == first_round[3]
    bits = torch.tensor([[(p >> (L - 1 - j)) & 1 for j in range(L)] for p in mine], dtype=torch.uint8).reshape(-1, L)
    gp, gf = gathered_pattern_vote(bits, torch.arange(a, b))
    out_q.put((rank, pattern.tolist(), freq, bit_votes.tolist(), frames, gp.tolist(), gf))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_vote_combine_world_size_2_gloo():
    L = 8
    # tie between 0x65 and 0x9A across the two ranks; 0x9A appears first globally (rank 0, frame 1)
    patterns = [0x01, 0x9A, 0x65, 0x65, 0x9A, 0x65, 0x9A, 0x02]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vote_worker, args=(r, 2, port, L, patterns, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want_p, want_f = _reference_vote(patterns, L)
    assert want_p == [int(c) for c in format(0x9A, "08b")]
    for rank, pattern, freq, bit_votes, frames, gp, gf in got:
        assert pattern == want_p and freq == want_f and frames == len(patterns)
        assert bit_votes == [sum((p >> (L - 1 - j)) & 1 for p in patterns) for j in range(L)]
        assert gp == want_p and gf == want_f


def _owned_vote_worker(rank, world, port, L, per_segment, out_q):
    from conftest import ROOT, PKG      # noqa: F401  (sets sys.path in the child)
    from b200wm.vote import SegmentVote
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_seg = len(per_segment)
    per = n_seg // world
    vote = SegmentVote(n_seg, L, "cpu", owned=(rank * per, per))

    def fill():
        state, first = vote._mine()
        order = 0
        for s in range(rank * per, (rank + 1) * per):           # whole segments live on one rank (config 4)
            for p in per_segment[s]:
                state["hist"][s - first, p] += 1
                state["first_seen"][s - first, p] = min(int(state["first_seen"][s - first, p]), order)
                state["seg_frames"][s - first] += 1
                for j in range(L):
                    state["bit_votes"][s - first, j] += (p >> (L - 1 - j)) & 1
                order += 1
    fill()
    vote.combine()
    once = [(None if r[0] is None else r[0].tolist(), r[1], r[2].tolist(), r[3]) for r in vote.result()]
    vote.reset()                                                # persistent state: next batch, overlapped combine
    fill()
    vote.combine(async_op=True)
    assert once == [(None if r[0] is None else r[0].tolist(), r[1], r[2].tolist(), r[3]) for r in vote.result()]
    out_q.put((rank, [(None if r[0] is None else r[0].tolist(), r[1], r[2].tolist(), r[3]) for r in vote.result()]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_owned_segment_vote_combine_world_size_2_gloo():
    """Whole segments dealt to ranks: the combine is one in-place all-gather of each rank's block and every
    rank ends with every segment's Counter result (ties -> first seen, empty segment -> None)."""
    L = 8
    per_segment = [[0x9A, 0x65, 0x65, 0x9A], [0x01] * 5 + [0x02] * 3, [], [0x65, 0x65, 0x10]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_owned_vote_worker, args=(r, 2, port, L, per_segment, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, res in got:
        for s, pats in enumerate(per_segment):
            pattern, freq, votes, frames = res[s]
            if not pats:
                assert pattern is None and freq is None and frames == 0
                continue
            want_p, want_f = _reference_vote(pats, L)
            assert pattern == want_p and freq == want_f and frames == len(pats), (rank, s)
            assert votes == [sum((p >> (L - 1 - j)) & 1 for p in pats) for j in range(L)]


def test_fingerprint_payload_schemes_and_decisions():
    """Host half of the fingerprint layer: the reference's payload schemes (tests/segment_mark_detect_hls.py:42-55,
    tests/mark_video_to_hls.py:27-43), their inverse (tests/detect_watermarks.py:145-172) and the copy sequence /
    fingerprint string (:404-425) - against the oracle's restatement."""
    from offmark_b200 import fingerprint as fp
    from oracle import payload as o_pay
    for seg in (0, 1, 15, 16, 17, 255, 256, 300):
        assert np.array_equal(fp.generate_payload_for_segment(seg), o_pay.payload_for_segment(seg))
        for copy in (0, 3, 15, 16):
            p = fp.generate_payload_for_segment(seg, copy)
            assert np.array_equal(p, o_pay.payload_for_segment_copy(seg, copy))
            assert fp.decode_watermark_pattern(p) == (seg % 16, copy % 16) == o_pay.segment_copy_from_pattern(p)
    assert fp.decode_watermark_pattern(None) == (None, None) and fp.decode_watermark_pattern([0, 1, 1]) == (None, None)
    results = [{"segment_number": 2, "detected_copy_index": 1}, {"segment_number": 0, "detected_copy_index": 3},
               {"segment_number": 1, "detected_copy_index": 0}]
    assert fp.copy_fingerprint(results) == ([3, 0, 1], "301")
    results[1]["detected_copy_index"] = None
    assert fp.copy_fingerprint(results) == ([None, 0, 1], None)


def test_segment_vote_and_packing_properties():
    """Property tests (hypothesis) of the host logic: SegmentVote.result() reproduces Counter.most_common(1) - ties go to
    the pattern seen first - for arbitrary pattern lists, orders of arrival and payload lengths; bit packing round-trips;
    the payload schemes invert."""
    from hypothesis import given, settings, strategies as st
    from b200wm import ops
    from b200wm.vote import SegmentVote
    from offmark_b200 import fingerprint as fp

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 6).flatmap(lambda L: st.tuples(st.just(L), st.lists(st.integers(0, (1 << L) - 1), min_size=1, max_size=40))),
           st.randoms(use_true_random=False))
    def vote_property(case, rnd):
        L, pats = case
        vote = SegmentVote(2, L, "cpu")
        order = list(range(len(pats)))
        rnd.shuffle(order)                                  # frames arrive in any order; first_seen keeps the frame index
        for i in order:
            _fill_vote_state(vote, [pats[i]], [1], [i])
        want_p, want_f = _reference_vote(pats, L)
        pattern, freq, votes, frames = vote.result()[1]
        assert pattern.tolist() == want_p and freq == want_f and frames == len(pats)
        assert vote.result()[0][0] is None
        vote.reset()
        assert vote.result()[1][0] is None and vote.result()[1][3] == 0
    vote_property()

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 4).flatmap(lambda r: st.integers(1, 200).flatmap(
        lambda n: st.lists(st.lists(st.integers(0, 1), min_size=n, max_size=n), min_size=r, max_size=r))))
    def packing_property(rows):
        packed, n = ops.pack_bits(np.array(rows))
        assert packed.shape == (len(rows), max(1, (n + 31) // 32)) and packed.dtype == torch.int32
        assert np.array_equal(ops.unpack_bits(packed, n), np.array(rows, dtype=np.uint8))
    packing_property()

    @settings(max_examples=100, deadline=None)
    @given(st.integers(0, 10 ** 6), st.integers(0, 10 ** 6))
    def payload_property(seg, copy):
        assert fp.decode_watermark_pattern(o_pay.payload_for_segment_copy(seg, copy)) == (seg % 16, copy % 16)
        assert int("".join(map(str, fp.generate_payload_for_segment(seg))), 2) == seg % 256
    payload_property()
    with pytest.raises(ValueError):
        ops.pack_bits(np.array([0, 2, 1]))


def test_batch_reader_and_writer_bookkeeping():
    """video/memory_io.py: the reference's read() / write() surface plus the optional batch protocol; reservations are
    committed in the order they were made and misuse is loud."""
    from offmark_b200.video.memory_io import BatchReader, BatchWriter
    frames = np.arange(5 * 4 * 6 * 3, dtype=np.uint8).reshape(5, 4, 6, 3)
    r = BatchReader(frames)
    assert (r.height, r.width) == (4, 6)
    assert np.array_equal(r.read(), frames[0]) and np.array_equal(r.read_batch(100), frames[1:]) and r.read() is None
    w = BatchWriter(4, frames[0].shape, pinned=False)
    a, b = w.reserve(2, frames[0].shape), w.reserve(2, frames[0].shape)
    assert a is not None and b is not None and w.reserve(1, frames[0].shape) is None and len(w.frames) == 0
    assert w.reserve(1, (4, 6, 1)) is None
    with pytest.raises(RuntimeError):
        w.write(frames[0])
    a[:] = frames[:2]
    w.commit(2)
    assert len(w.frames) == 2 and np.array_equal(w.frames, frames[:2])
    with pytest.raises(RuntimeError):
        w.commit(3)
    w.rewind()
    w.write(frames[4])
    assert len(w.frames) == 1 and np.array_equal(w.frames[0], frames[4])


def test_gathered_batches_cut_at_size_and_at_shape_changes():
    """offmark_b200/_frames.py:gathered_batches - the reference's read() loop (video/embedder.py:19-27) gathered into
    one-shape batches for the batched drivers."""
    from offmark_b200._frames import gathered_batches
    from offmark_b200.video.memory_io import ArrayReader
    shapes = [(4, 6)] * 4 + [(6, 4)] * 2 + [(4, 6)] * 5
    frames = [np.full((h, w, 3), k, dtype=np.int64) for k, (h, w) in enumerate(shapes)]
    got = list(gathered_batches(ArrayReader(frames), 3))
    assert [len(g) for g in got] == [3, 1, 2, 3, 2]
    flat = [f for g in got for f in g]
    assert all(f.dtype == np.uint8 and f.flags["C_CONTIGUOUS"] for f in flat)
    assert [int(f[0, 0, 0]) for f in flat] == list(range(len(frames))) and [f.shape[:2] for f in flat] == shapes
    assert list(gathered_batches(ArrayReader(frames[:1]), 8))[0][0].shape == (4, 6, 3)

    class Empty:
        def read(self):
            return None
    assert list(gathered_batches(Empty(), 4)) == []
