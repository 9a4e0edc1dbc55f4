"""GPU: vote kernels vs the oracle's DeShuffler / Counter restatement."""
import json
import os
from collections import Counter

import numpy as np
import pytest
import torch

from oracle import payload as o_pay

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bits_with_counts(length, n, counts):
    bits = np.zeros(n)
    for i, c in enumerate(counts):
        idx = np.arange(i, n, length)
        bits[idx[:c]] = 1
    return bits


def test_counts_and_finish_match_reference_cases(golden_dir):
    from b200wm import ops
    with open(os.path.join(golden_dir, "payload.json")) as f:
        cases = json.load(f)["degenerate_cases"]
    for case in cases:
        length, n, key = case["length"], case["n"], case["key"]
        bits = _bits_with_counts(length, n, case["counts"])
        packed, _ = ops.pack_bits(bits, device=DEV)
        counts = ops.vote_counts(packed, n, length)
        assert counts[0].cpu().tolist() == case["counts"]
        perm = torch.from_numpy(o_pay.permutation(length, key).astype(np.int32)).to(DEV)
        patterns, word = ops.vote_finish(counts, n, perm)
        assert patterns[0].cpu().tolist() == case["pattern"], case
        assert int(word[0]) == int("".join(map(str, case["pattern"])), 2)


def test_exact_ties_follow_float64_expression():
    """Ties are where an integer majority rule and the reference's float64 threshold disagree."""
    from b200wm import ops
    rng = np.random.RandomState(1)
    n, length = 32400, 8
    rows, want = [], []
    for _ in range(200):
        base = rng.randint(0, 4051)
        counts = np.clip(base + rng.randint(-1, 2, length), 0, 4050)
        bits = _bits_with_counts(length, n, counts)
        rows.append(counts)
        want.append(o_pay.degenerate(bits.reshape(1, -1), length, 0))
    counts_t = torch.tensor(np.array(rows), dtype=torch.int32, device=DEV)
    perm = torch.from_numpy(o_pay.permutation(length, 0).astype(np.int32)).to(DEV)
    patterns, _ = ops.vote_finish(counts_t, n, perm)
    assert np.array_equal(patterns.cpu().numpy(), np.array(want))


@pytest.mark.parametrize("length,n", [(8, 32400), (441, 32400), (5, 97), (32, 1200), (3, 129600), (9000, 32400), (16, 7)])
def test_general_payload_lengths(length, n):
    from b200wm import ops
    rng = np.random.RandomState(length)
    bits = (rng.rand(4, n) < 0.4).astype(np.uint8)
    packed, _ = ops.pack_bits(bits, device=DEV)
    counts = ops.vote_counts(packed, n, length)
    want_counts = np.array([[int(b[i::length].sum()) for i in range(length)] for b in bits])
    assert np.array_equal(counts.cpu().numpy(), want_counts)
    perm = torch.from_numpy(o_pay.permutation(length, 3).astype(np.int32)).to(DEV)
    patterns, _ = ops.vote_finish(counts, n, perm)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.array([o_pay.degenerate(b.reshape(1, -1).astype(np.float64), length, 3) for b in bits])
    assert np.array_equal(patterns.cpu().numpy(), want)


def test_pattern_hist_reproduces_counter_mode():
    from b200wm.vote import SegmentVote
    rng = np.random.RandomState(0)
    L, S, n = 8, 7, 420
    seg = rng.randint(0, S - 1, n).astype(np.int32)           # last segment stays empty
    pats = rng.choice([0x65, 0x9A, 0x11, 0xF0], size=n, p=[0.3, 0.3, 0.2, 0.2])
    vote = SegmentVote(S, L, DEV)
    # two calls with an order offset, as a rank processing two batches would do
    half = n // 2
    for a, b in ((0, half), (half, n)):
        vote.add(torch.tensor(pats[a:b], dtype=torch.int64, device=DEV),
                 frame_segment=torch.tensor(seg[a:b], device=DEV), order_offset=a)
    res = vote.result()
    for s in range(S):
        mine = [format(p, "08b") for p, sg in zip(pats, seg) if sg == s]
        if not mine:
            assert res[s][0] is None
            continue
        best, count = Counter(mine).most_common(1)[0]
        assert "".join(map(str, res[s][0])) == best and res[s][1] == count / len(mine) and res[s][3] == len(mine)
        assert res[s][2].tolist() == [sum(int(m[j]) for m in mine) for j in range(L)]


def test_owned_block_layout_equals_general_layout():
    """SegmentVote(owned=...) (rank-major state, global segment numbers in, one all-gather to combine) gives the
    same per-segment results as the general layout; here as the second of two equal blocks of a 2-rank job."""
    from b200wm.vote import SegmentVote
    rng = np.random.RandomState(1)
    L, per, n = 8, 5, 300
    seg = (per + rng.randint(0, per, n)).astype(np.int32)      # this "rank" owns segments 5..9 of 10
    pats = rng.choice([0x65, 0x9A, 0x11], size=n, p=[0.4, 0.4, 0.2])
    packed = torch.tensor(pats, dtype=torch.int64, device=DEV)
    fs = torch.tensor(seg, device=DEV)
    general = SegmentVote(2 * per, L, DEV).add(packed, frame_segment=fs, order_offset=1000).result()
    owned = SegmentVote(2 * per, L, DEV, owned=(per, per)).add(packed, frame_segment=fs, order_offset=1000).result()
    for a, b in zip(general, owned):
        assert (a[0] is None) == (b[0] is None) and a[1] == b[1] and a[3] == b[3] and np.array_equal(a[2], b[2])
        if a[0] is not None:
            assert np.array_equal(a[0], b[0])
    with pytest.raises(ValueError):
        SegmentVote(10, L, DEV, owned=(3, 5))


def test_pattern_hist_publish_equals_pattern_hist_on_one_gpu():
    """b200wm_pattern_hist_publish (histogram fused with the NVLink exchange of the finished block) with a world of one
    rank: same state as b200wm_pattern_hist, counters re-armed, and the wait kernel returns at once.  The multi-GPU
    exchange itself is checked against the NCCL all-gather by scripts/publish_check.py under torchrun (2 and 8 GPUs)."""
    from b200wm import ops
    dev = torch.device("cuda:0")
    L, n_seg = 8, 7
    gen = torch.Generator(device=dev).manual_seed(3)
    for n in (0, 1, 300, 5000):
        packed = torch.randint(0, 1 << L, (n,), device=dev, generator=gen, dtype=torch.int64)
        seg = torch.randint(0, n_seg, (n,), device=dev, generator=gen, dtype=torch.int32)
        want = ops.pattern_hist(packed, L, n_seg, seg, None, 11) if n else None
        bins = 1 << L
        block_len = (n_seg * (2 * bins + L + 1) + 3) // 4 * 4
        buf = torch.zeros(block_len + 4, dtype=torch.int32, device=dev)
        a, b, c = n_seg * bins, n_seg * (bins + L), n_seg * (bins + L + 1)
        state = {"hist": buf[:a].view(n_seg, bins), "bit_votes": buf[a:b].view(n_seg, L), "seg_frames": buf[b:c],
                 "first_seen": buf[c:c + n_seg * bins].view(n_seg, bins)}
        state["first_seen"].fill_(ops.INT32_MAX)
        peers = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
        ticket, status = torch.zeros(2, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
        for epoch in (1, 2):                     # twice: the counters must have been re-armed by the first launch
            ops.pattern_hist_publish(packed, L, n_seg, state, seg, None, 11, peers.data_ptr(), block_len, 1, 0, epoch, ticket, status)
            ops.vote_exchange_wait(peers.data_ptr(), block_len, 1, 0, epoch, status)
        torch.cuda.synchronize()
        assert ticket.tolist() == [0, 0] and status.item() == 0
        if n:
            assert torch.equal(state["hist"], 2 * want["hist"]) and torch.equal(state["first_seen"], want["first_seen"])
            assert torch.equal(state["bit_votes"], 2 * want["bit_votes"]) and torch.equal(state["seg_frames"], 2 * want["seg_frames"])
        else:
            assert not state["hist"].any()


def test_symmetric_segment_vote_in_a_one_rank_group():
    """SegmentVote(symmetric=True) - state in peer-mapped memory, histogram kernel fused with the exchange - inside a
    process group of one rank: same results as the plain state, across reset and re-use, and for a batch longer than one
    fused launch may take (the head then goes through the plain histogram kernel)."""
    import socket
    import torch.distributed as dist
    from b200wm.vote import SegmentVote
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
    try:
        L, n_seg = 8, 5
        try:
            fused = SegmentVote(n_seg, L, dev, owned=(0, n_seg), symmetric=True)
        except Exception as exc:                 # no peer-mappable memory on this box
            pytest.skip(f"symmetric memory unavailable: {exc}")
        plain = SegmentVote(n_seg, L, dev)
        gen = torch.Generator(device=dev).manual_seed(9)
        for n in (700, 40000, 3):                # 40000 > 128 * 256 frames: head + fused tail
            packed = torch.randint(0, 1 << L, (n,), device=dev, generator=gen, dtype=torch.int64)
            seg = torch.randint(0, n_seg, (n,), device=dev, generator=gen, dtype=torch.int32)
            a = plain.reset().add(packed, frame_segment=seg, order_offset=5).combine().result()
            b = fused.reset().add(packed, frame_segment=seg, order_offset=5).combine().result()
            for x, y in zip(a, b):
                assert (x[0] is None and y[0] is None) or (x[0].tolist() == y[0].tolist() and x[1] == y[1])
                assert x[2].tolist() == y[2].tolist() and x[3] == y[3]
            assert torch.equal(plain.first_seen.reshape(-1), fused.first_seen.reshape(-1))
    finally:
        dist.destroy_process_group()


def test_pattern_hist_ignores_what_is_out_of_range_and_rejects_wrong_types():
    """The histogram kernels read 8 bytes per frame: anything but int64 patterns is refused by the wrapper, and a frame
    whose segment or pattern is out of range is skipped by the kernel instead of writing outside the tables."""
    from b200wm import ops
    from b200wm.vote import SegmentVote
    L, n_seg = 8, 2
    packed = torch.tensor([3, 300, 5, 3, -1], dtype=torch.int64, device=DEV)          # 300 and -1 are not 8-bit patterns
    seg = torch.tensor([0, 0, 9, 1, 1], dtype=torch.int32, device=DEV)                # 9 is not a segment
    st = ops.pattern_hist(packed, L, n_seg, seg)
    torch.cuda.synchronize()
    assert st["seg_frames"].tolist() == [1, 1] and int(st["hist"].sum()) == 2
    assert int(st["hist"][0, 3]) == 1 and int(st["hist"][1, 3]) == 1
    assert st["first_seen"][0, 3].item() == 0 and st["first_seen"][1, 3].item() == 3
    res = SegmentVote(n_seg, L, DEV).add(packed, frame_segment=seg.to(torch.int64)).result()       # coerced to int32
    assert [r[1] for r in res] == [1, 1]
    with pytest.raises(ValueError):
        ops.pattern_hist(packed.to(torch.int32), L, n_seg, seg)
    with pytest.raises(ValueError):
        ops.pattern_hist(packed, L, n_seg, seg.to(torch.int64))
    with pytest.raises(ValueError):
        ops.pattern_hist(packed, L, n_seg, seg[:3].contiguous())
    with pytest.raises(ValueError):
        ops.pattern_hist(packed.cpu(), L, n_seg, None)
    with pytest.raises(ValueError):
        ops.pattern_hist(packed, 17, n_seg, seg)


def test_a_table_that_is_not_a_permutation_never_writes_outside_the_patterns():
    """``perm`` lives on the GPU for b200wm_vote_finish, so its entries are guarded in the kernel (an entry that is not
    a payload position is skipped); the host-buffer entry points see the table and refuse it."""
    from b200wm import ops
    L, n = 8, 4
    counts = torch.full((n, L), 3, dtype=torch.int32, device=DEV)
    counts[:, 2] = 9
    perm = torch.tensor([0, 1, 2, 3, 99, -5, 6, 7], dtype=torch.int32, device=DEV)
    patterns, _ = ops.vote_finish(counts, 10 * L, perm)
    good, _ = ops.vote_finish(counts, 10 * L, torch.arange(L, dtype=torch.int32, device=DEV))
    torch.cuda.synchronize()
    keep = [0, 1, 2, 3, 6, 7]
    assert torch.equal(patterns[:, keep], good[:, keep]) and good[:, 2].tolist() == [1] * n
    planes = np.zeros((2, 64, 64), dtype=np.uint8)
    with pytest.raises(ValueError):
        ops.dwtsvd_detect_host(planes, np.array([0, 1, 2, 3, 4, 5, 6, 8], dtype=np.int32))
    rows = ops.pack_bits(np.zeros((1, 64), dtype=np.int64))[0].cpu().contiguous()
    with pytest.raises(ValueError):
        ops.dwtsvd_mark_verify_host(planes, np.empty_like(planes), rows, np.array([0, 1, 2, 3, 4, 5, 6, -1], dtype=np.int32), wm_len=64)
    with pytest.raises(ValueError):
        ops.vote_counts(torch.zeros((2, 2), dtype=torch.int64, device=DEV), 64, 8)
    with pytest.raises(ValueError):
        ops.vote_counts(torch.zeros((2, 2), dtype=torch.int32, device=DEV), 65, 8)
