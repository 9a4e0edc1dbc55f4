"""GPU: the reference's driver flows running on the drop-in plugin classes
(tests/mark.py + tests/detect.py, tests/test.py combos 0:0 / 0:3 / 1:0, HLS pattern collector)."""
import os

import numpy as np
import pytest
import torch

from oracle import bracket, dwt_dct_svd as o_svd, payload as o_pay
from parity import PAYLOAD, KEY

pytestmark = pytest.mark.gpu


def test_mark_py_and_detect_py_flow_on_in_mp4_frames(golden_dir):
    """mark.py:18-40 then detect.py:17-31 with the in-memory reader/writer."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.extractor import Extractor
    from offmark_b200.video.memory_io import ArrayReader, ArrayWriter

    g = np.load(os.path.join(golden_dir, "in_mp4_frames.npz"))
    frames = [g[f"rgb_{i}"] for i in g["picks"]]
    r, w = ArrayReader(frames), ArrayWriter()
    frame_embedder = DwtDctSvdEncoder()
    capacity = frame_embedder.wm_capacity((r.height, r.width, 3))
    wm = Shuffler(key=KEY).generate_wm(PAYLOAD, capacity)
    assert np.array_equal(wm.astype(np.uint8), g["wm"])
    frame_embedder.read_wm(wm)
    Embedder(r, frame_embedder, w).start()
    assert len(w.frames) == len(frames)
    for i, src, marked in zip(g["picks"], frames, w.frames):
        ref_marked = (src.astype(np.int16) + g[f"marked_minus_src_{i}"]).astype(np.uint8)
        d = np.abs(marked.astype(np.int16) - ref_marked)
        assert d.max() <= 1 and (d > 0).mean() < 2e-3, (d.max(), (d > 0).mean())

    degenerator = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
    extractor = Extractor(ArrayReader(w.frames), DwtDctSvdDecoder(), degenerator)
    for i, marked in zip(g["picks"], w.frames):
        assert np.array_equal(extractor.check_frame(marked), g["patterns_all_frames"][i])
        # and on the reference's own marked frame: the per-frame pattern the reference reported
        ref_marked = (g[f"rgb_{i}"].astype(np.int16) + g[f"marked_minus_src_{i}"]).astype(np.uint8)
        assert np.array_equal(extractor.check_frame(ref_marked), g["patterns_all_frames"][i])
    extractor.start()


def test_test_py_flow_numpy_in_place_semantics(golden_dir):
    """tests/test.py:85-121 with numpy frames: encode mutates and returns its argument."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.embed.dct_encoder import DctEncoder
    from offmark_b200.extract.dct_decoder import DctDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler

    frame = np.load(os.path.join(golden_dir, "frame63_crop.npz"))["bgr"]
    for enc, dec, o_dec in ((DwtDctSvdEncoder(), DwtDctSvdDecoder(), o_svd.decode),
                            (DctEncoder(), DctDecoder(), None)):
        yuv = bracket.to_yuv(frame)
        wm = Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(yuv.shape))
        enc.read_wm(wm)
        before = yuv.copy()
        out = enc.encode(yuv)
        assert out is yuv and not np.array_equal(yuv, before)
        assert np.array_equal(yuv[:, :, 0], before[:, :, 0])
        marked = bracket.from_yuv(yuv.copy())
        decoded = dec.decode(bracket.to_yuv(marked))
        assert isinstance(decoded, np.ndarray) and decoded.dtype == np.float64
        assert decoded.shape == (1, frame.shape[0] * frame.shape[1] // 64)
        ret = DeShuffler(key=KEY).set_shape(PAYLOAD.shape).degenerate(decoded)
        assert ret.dtype == np.uint8 and np.array_equal(ret, PAYLOAD)
        # a plain ndarray (no device side-car) goes through the same kernels
        assert np.array_equal(DeShuffler(key=KEY).set_shape(PAYLOAD.shape).degenerate(np.array(decoded)), PAYLOAD)
        if o_dec is not None:
            assert np.array_equal(o_pay.degenerate(o_dec(bracket.to_yuv(marked)), 8, KEY), PAYLOAD)


def test_grayscale_payload_roundtrip():
    """tests/test.py combo 1:0 (GrayScale generator + DwtDctSvd coder) with a 21x21 image payload."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.grayscale import GrayScale
    from offmark_b200.degenerator.de_grayscale import DeGrayScale
    from oracle import synth
    img = (np.random.RandomState(4).rand(21, 21) > 0.5).astype(np.uint8) * 255
    frame = synth.random_bgr(1080, 1920, 8)
    yuv = bracket.to_yuv(frame)
    enc = DwtDctSvdEncoder()
    enc.read_wm(GrayScale(key=KEY).generate_wm(img, enc.wm_capacity(yuv.shape)))
    enc.encode(yuv)
    back = DeGrayScale(key=KEY).set_shape(img.shape).degenerate(DwtDctSvdDecoder().decode(yuv))
    assert back.shape == img.shape and np.array_equal(back, img)


def test_hls_pattern_collector_flow():
    """tests/segment_mark_detect_hls.py: per-segment payload = 8-bit segment number (:42-55),
    per-frame patterns, Counter mode and frequency (:126-155) - batched on the GPU."""
    from b200wm import ops
    from b200wm.vote import SegmentVote
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    from oracle import synth
    dev = torch.device("cuda:0")
    S, F, h, w = 5, 12, 240, 320
    planes = torch.from_numpy(np.stack([synth.luma_plane_u8(h, w, f, 31) for f in range(S * F)])).to(dev)
    seg_ids = [3, 77, 130, 254, 257]
    payloads = [o_pay.payload_for_segment(s) for s in seg_ids]
    enc, dec = DwtDctSvdEncoder(), DwtDctSvdDecoder()
    rows = np.stack([Shuffler(key=KEY).generate_wm(p, enc.wm_capacity((h, w, 3)))[0] for p in payloads])
    frame_seg = torch.arange(S * F, device=dev, dtype=torch.int32) // F
    enc.read_wm(rows[:1])
    enc.encode_planes(planes, wm_rows=rows, frame_wm_row=frame_seg)
    raw, counts = dec.decode_planes(planes, payload_len=8)
    deg = DeShuffler(key=KEY).set_shape((8,))
    patterns, packed = deg.degenerate_counts(counts, h * w // 64)
    vote = SegmentVote(S, 8, dev).add(packed, frame_segment=frame_seg).combine()
    for s, (pattern, freq, bit_votes, frames) in enumerate(vote.result()):
        assert np.array_equal(pattern, payloads[s]) and freq == 1.0 and frames == F
        assert bit_votes.tolist() == (payloads[s] * F).tolist()
    # the same frames through the oracle's per-frame path
    host = planes.cpu().numpy()
    for f in (0, F + 1, 4 * F + 3):
        bits = o_svd.extract_plane(host[f])
        assert np.array_equal(o_pay.degenerate(bits, 8, KEY), patterns[f].cpu().numpy())


def test_host_buffer_entry_points_match_device_path():
    """b200wm_dwtsvd_mark_host / _detect_host (frames in host memory, chunked double-buffered streaming
    inside the library) give exactly what the device-resident calls give."""
    from b200wm import ops
    from oracle import synth
    n, h, w = 7, 240, 320
    planes = np.stack([synth.luma_plane_u8(h, w, f, 9) for f in range(n)])
    payloads = [o_pay.payload_for_segment(s) for s in (9, 100)]
    rows = np.stack([o_pay.generate_wm(p, (1, h * w // 64), KEY)[0] for p in payloads])
    frame_row = np.array([0, 0, 0, 1, 1, 1, 1], dtype=np.int32)
    # I420-like host layout: planes are strided views of a bigger pinned buffer
    frame_bytes = h * w * 3 // 2
    host = torch.zeros((n, frame_bytes), dtype=torch.uint8).pin_memory()
    src = host.as_strided((n, h, w), (frame_bytes, w, 1))
    src.copy_(torch.from_numpy(planes))
    dst = torch.empty_like(host).as_strided((n, h, w), (frame_bytes, w, 1))
    for chunk in (0, 1, 3):
        ops.dwtsvd_mark_host(src, dst, rows, frame_wm_row=frame_row, chunk_frames=chunk)
        dev = torch.from_numpy(planes).cuda()
        packed, nb = ops.pack_bits(rows, device="cuda")
        ops.dwtsvd_embed_(dev, packed, nb, frame_wm_row=torch.from_numpy(frame_row).cuda())
        assert np.array_equal(dst.numpy(), dev.cpu().numpy())
        patterns, raw = ops.dwtsvd_detect_host(dst, o_pay.permutation(8, KEY), chunk_frames=chunk, want_raw_bits=True)
        raw_dev, counts = ops.dwtsvd_extract(dev, payload_len=8)
        assert np.array_equal(raw, raw_dev.cpu().numpy().view(np.uint32))
        for f in range(n):
            assert np.array_equal(patterns[f], payloads[frame_row[f]])
    # in place on the host buffer, pageable memory
    pageable = planes.copy()
    ops.dwtsvd_mark_host(pageable, pageable, rows, frame_wm_row=frame_row)
    assert np.array_equal(pageable, dst.numpy())


def test_embedder_batched_mode_equals_per_frame_mode(golden_dir):
    """Embedder(batch_frames=N) (optional extension: one upload / launch / download per N frames) writes the
    very bytes the reference-shaped per-frame loop writes, in order, including a ragged last batch."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.memory_io import ArrayReader, ArrayWriter
    from oracle import synth
    frames = [synth.random_bgr(64, 96, s) for s in range(17)]            # 6 batches of 3: the three-slot pipeline wraps

    def run(batch, clip=frames):
        enc = DwtDctSvdEncoder()
        enc.read_wm(Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(clip[0].shape)))
        w = ArrayWriter()
        Embedder(ArrayReader([f.copy() for f in clip]), enc, w, batch_frames=batch).start()
        return w.frames
    one, many = run(1), run(3)
    assert len(one) == len(many) == len(frames)
    for a, b in zip(one, many):
        assert a.dtype == np.uint8 and np.array_equal(a, b)
    # frames above 1 MiB are gathered by the copy threads; 22 frames in batches of 4 (ragged end)
    big = [synth.random_bgr(512, 704, 100 + s) for s in range(22)]
    one, many = run(1, big), run(4, big)
    assert len(one) == len(many) == len(big)
    for a, b in zip(one, many):
        assert np.array_equal(a, b)


def test_embedder_batched_mode_cuts_batches_where_the_frame_shape_changes():
    """A batch is one launch, so it holds one shape: a clip whose geometry changes mid-stream is cut there (the reference's
    loop has no such restriction, embedder.py:19-27) and the marked frames still equal the per-frame path's, in order."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.memory_io import ArrayReader, ArrayWriter
    from offmark_b200._frames import gathered_batches
    from oracle import synth
    shapes = [(64, 96)] * 4 + [(96, 64)] * 2 + [(64, 96)] * 5
    frames = [synth.random_bgr(h, w, s) for s, (h, w) in enumerate(shapes)]
    assert [len(g) for g in gathered_batches(ArrayReader(frames), 3)] == [3, 1, 2, 3, 2]

    def run(batch):
        w = ArrayWriter()

        class PerShapeEncoder(DwtDctSvdEncoder):                 # the watermark is tiled to the frame's capacity
            def mark_rgb8(self, f):
                self.read_wm(Shuffler(key=KEY).generate_wm(PAYLOAD, self.wm_capacity(f.shape[-3:])))
                return super().mark_rgb8(f)
        Embedder(ArrayReader([f.copy() for f in frames]), PerShapeEncoder(), w, batch_frames=batch).start()
        return w.frames
    one, many = run(1), run(3)
    assert len(one) == len(many) == len(frames)
    for a, b in zip(one, many):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_extractor_batched_mode_equals_per_frame_mode():
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.extractor import Extractor
    from offmark_b200.video.memory_io import ArrayReader, ArrayWriter
    from oracle import synth
    frames = [synth.random_bgr(128, 192, s) for s in range(9)]
    w = ArrayWriter()
    for part, payload in ((frames[:4], PAYLOAD), (frames[4:], 1 - PAYLOAD)):     # two payloads: the order of the patterns is visible
        enc = DwtDctSvdEncoder()
        enc.read_wm(Shuffler(key=KEY).generate_wm(payload, enc.wm_capacity(frames[0].shape)))
        Embedder(ArrayReader(part), enc, w, batch_frames=4).start()

    def run(batch):
        ex = Extractor(ArrayReader(w.frames), DwtDctSvdDecoder(), DeShuffler(key=KEY).set_shape((8,)), batch_frames=batch)
        ex.start()
        return ex.patterns
    one, many = run(1), run(2)
    assert len(one) == len(many) == len(frames)
    for k, (a, b) in enumerate(zip(one, many)):
        assert np.array_equal(a, b) and np.array_equal(a, PAYLOAD if k < 4 else 1 - PAYLOAD)


def test_fingerprint_layer_copy_sequence_roundtrip():
    """tests/mark_video_to_hls.py + tests/detect_watermarks.py on the GPU: N marked copies per segment from one
    read, a player picks one copy per segment, detection recovers the copy sequence - with the payload map
    and, without it, by decoding the pattern; payload scheme and JSON shapes as in the reference."""
    from offmark_b200 import fingerprint as fp
    from oracle import synth
    dev = torch.device("cuda:0")
    n_seg, per, n_copies, h, w = 4, 6, 3, 240, 320
    seg_numbers = [0, 1, 2, 17]                     # 17 % 16 == 1: the 4-bit scheme wraps (mark_video_to_hls.py:38)
    frame_seg = [seg_numbers[f // per] for f in range(n_seg * per)]
    planes = torch.from_numpy(np.stack([synth.luma_plane_u8(h, w, f, 77) for f in range(n_seg * per)])).to(dev)
    copies, payloads, seg_copies = fp.mark_segment_copies(planes, frame_seg, n_copies)
    assert copies.shape == (n_copies, n_seg * per, h, w)
    assert payloads["2_1"] == o_pay.payload_for_segment_copy(2, 1).tolist() == [0, 0, 1, 0, 0, 0, 0, 1]
    assert seg_copies["17"][2] == {"file": "marked_seg17_copy2.mp4", "payload": o_pay.payload_for_segment_copy(17, 2).tolist(), "copy_index": 2}
    assert np.array_equal(fp.generate_payload_for_segment(300), o_pay.payload_for_segment(300))
    chosen = [2, 0, 1, 2]                           # the copy this player received for each segment
    stream = torch.cat([copies[chosen[i], i * per:(i + 1) * per] for i in range(n_seg)])
    with_map = fp.detect_segment_copies(stream, frame_seg, segment_payloads=payloads)
    assert [r["detected_copy_index"] for r in with_map] == [2, 0, 1, 2]
    assert all(r["success"] and r["match_frequency"] == 1.0 for r in with_map)
    assert fp.copy_fingerprint(with_map) == ([2, 0, 1, 2], "2012")
    without = fp.detect_segment_copies(stream, frame_seg)
    assert [r["detected_copy_index"] for r in without] == [2, 0, 1, 2]
    assert fp.decode_watermark_pattern(o_pay.payload_for_segment_copy(5, 9)) == o_pay.segment_copy_from_pattern(o_pay.payload_for_segment_copy(5, 9)) == (5, 9)
    # a payload map that does not list the received copies identifies nothing (detect_watermarks.py:339-347)
    other = {k: v for k, v in payloads.items() if k.endswith("_1")}
    partial = fp.detect_segment_copies(stream, frame_seg, segment_payloads=other)
    assert [r["detected_copy_index"] for r in partial] == [None, None, 1, None]


def test_mark_verify_host_equals_mark_then_detect():
    """b200wm_dwtsvd_mark_verify_host (one upload, embed, extract + vote on the resident marked chunk, one download:
    the reference's mark-then-verify step, tests/mark_video_to_hls.py:356-399) returns the very bytes of
    b200wm_dwtsvd_mark_host and the very patterns / raw bits of b200wm_dwtsvd_detect_host on them."""
    from b200wm import ops
    from oracle import synth
    n, h, w = 9, 240, 320
    planes = np.stack([synth.luma_plane_u8(h, w, f, 19) for f in range(n)])
    payloads = [o_pay.payload_for_segment(s) for s in (9, 100, 201)]
    rows = np.stack([o_pay.generate_wm(p, (1, h * w // 64), KEY)[0] for p in payloads])
    frame_row = np.array([0, 0, 0, 1, 1, 1, 2, 2, 2], dtype=np.int32)
    src = torch.from_numpy(planes).pin_memory()
    perm = o_pay.permutation(8, KEY)
    for chunk in (0, 1, 4):
        two_a, one = torch.empty_like(src), torch.empty_like(src)
        ops.dwtsvd_mark_host(src, two_a, rows, frame_wm_row=frame_row, chunk_frames=chunk)
        pat_two, raw_two = ops.dwtsvd_detect_host(two_a, perm, chunk_frames=chunk, want_raw_bits=True)
        pat_one, raw_one = ops.dwtsvd_mark_verify_host(src, one, rows, perm, frame_wm_row=frame_row, chunk_frames=chunk,
                                                       want_raw_bits=True)
        assert np.array_equal(one.numpy(), two_a.numpy())
        assert np.array_equal(pat_one, pat_two) and np.array_equal(raw_one, raw_two)
        for f in range(n):
            assert np.array_equal(pat_one[f], payloads[frame_row[f]])
    # against the oracle on one frame
    want = o_svd.embed_plane_u8(planes[4], rows[1])
    assert np.abs(one.numpy()[4].astype(np.int16) - want).max() <= 1
    with pytest.raises(ValueError):                 # a row index outside the table is an error on the host path
        ops.dwtsvd_mark_verify_host(src, one, rows, perm, frame_wm_row=np.full(n, 3, dtype=np.int32))
    with pytest.raises(ValueError):
        ops.dwtsvd_mark_host(src, one, rows, frame_wm_row=np.full(n, -1, dtype=np.int32))


def test_degenerate_votes_on_the_array_it_is_given():
    """The reference's DeShuffler.degenerate works on whatever array it receives (de_shuffler.py:14-22): an inverted,
    sliced or edited copy of decode()'s result must give the vote of THAT array, not of the decoder's cached bits."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    from oracle import synth
    yuv = bracket.to_yuv(synth.random_bgr(240, 320, 5))
    enc = DwtDctSvdEncoder()
    enc.read_wm(Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(yuv.shape)))
    enc.encode(yuv)
    bits = DwtDctSvdDecoder().decode(yuv)
    assert type(bits) is np.ndarray and bits.dtype == np.float64
    deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
    assert np.array_equal(deg.degenerate(bits), PAYLOAD)
    assert np.array_equal(deg.degenerate(1 - bits), 1 - PAYLOAD)
    assert np.array_equal(deg.degenerate(1 - bits), o_pay.degenerate(1 - bits, 8, KEY))
    half = bits[:, :bits.shape[1] // 2 // 8 * 8]
    assert np.array_equal(deg.degenerate(half), o_pay.degenerate(half, 8, KEY))
    edited = bits.copy()
    edited[0, 6::8] = 1 - edited[0, 6::8]           # position 6 -> payload index perm[6]
    want = o_pay.degenerate(edited, 8, KEY)
    assert not np.array_equal(want, PAYLOAD) and np.array_equal(deg.degenerate(edited), want)
    bits[0, 6::8] = 1 - bits[0, 6::8]               # the same edit in place
    assert np.array_equal(deg.degenerate(bits), want)


def test_extractor_batched_mode_with_degrayscale_formats_like_per_frame_mode():
    """Extractor(batch_frames=N) with a DeGrayScale degenerator logs the 0/255 image of payload_shape that the
    per-frame path (and the reference, de_grayscale.py:15-23) returns, not flat 0/1 vectors."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.grayscale import GrayScale
    from offmark_b200.degenerator.de_grayscale import DeGrayScale
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.extractor import Extractor
    from offmark_b200.video.memory_io import ArrayReader, ArrayWriter
    from oracle import synth
    img = (np.random.RandomState(5).rand(4, 4) > 0.5).astype(np.uint8) * 255
    frames = [synth.random_bgr(240, 320, s) for s in range(5)]
    enc = DwtDctSvdEncoder()
    enc.read_wm(GrayScale(key=KEY).generate_wm(img, enc.wm_capacity(frames[0].shape)))
    w = ArrayWriter()
    Embedder(ArrayReader(frames), enc, w, batch_frames=2).start()

    def run(batch):
        ex = Extractor(ArrayReader(w.frames), DwtDctSvdDecoder(), DeGrayScale(key=KEY).set_shape(img.shape), batch_frames=batch)
        ex.start()
        return ex.patterns
    one, many = run(1), run(4)
    assert len(one) == len(many) == len(frames)
    for a, b in zip(one, many):
        assert a.shape == img.shape == b.shape and a.dtype == b.dtype == np.uint8
        assert np.array_equal(a, b) and np.array_equal(a, img)


def test_argument_checks_of_the_wrappers():
    """a CPU watermark tensor or an out-of-range row must be a ValueError, never a device fault; planes
    with a dimension below 8 have raw-bit words but no tile."""
    from b200wm import ops
    dev = torch.device("cuda:0")
    frames = torch.zeros((2, 64, 64, 3), dtype=torch.uint8, device=dev)
    planes = torch.zeros((2, 64, 64), dtype=torch.uint8, device=dev)
    packed_cpu, n = ops.pack_bits(np.zeros((2, 64), dtype=np.int64))
    packed, _ = ops.pack_bits(np.zeros((2, 64), dtype=np.int64), device=dev)
    masks = ops.dct8_masks(planes)
    for call in (lambda wm, **k: ops.dwtsvd_embed_rgb8_(frames, wm, n, **k),
                 lambda wm, **k: ops.dct8_embed_(planes, masks, wm, n, **k),
                 lambda wm, **k: ops.dwtsvd_embed_(planes, wm, n, **k)):
        with pytest.raises(ValueError):
            call(packed_cpu)
        with pytest.raises(ValueError):
            call(packed.to(torch.int64))
        bad = torch.tensor([0, 2], dtype=torch.int32, device=dev)
        with pytest.raises(ValueError):
            call(packed, frame_wm_row=bad, validate_rows=True)
        call(packed, frame_wm_row=bad)                       # unvalidated: the device clamps the row, no fault
        torch.cuda.synchronize()
    for h, w in ((4, 64), (64, 4), (4, 512)):
        t = torch.full((3, h, w), 200, dtype=torch.uint8, device=dev)
        for path in (0, 1):
            ops.set_path(path)
            raw, counts = ops.dwtsvd_extract(t, payload_len=8)
            assert raw.shape == (3, (h * w // 64 + 31) // 32)
            assert not raw.any().item() and not counts.any().item()
        ops.set_path(0)
    torch.cuda.synchronize()


def test_batch_reader_writer_protocol_equals_per_frame_flow():
    """Optional batch protocol of the drivers (video/memory_io.py: BatchReader.read_batch, BatchWriter.reserve / commit):
    frames go up and come down as whole batches from / into the reader's and writer's own (pinned) memory; the marked
    bytes and the logged patterns equal the reference-shaped per-frame loop's, including a ragged last batch, and the
    plain read() / write() of the same objects still work."""
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.generator.shuffler import Shuffler
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    from offmark_b200.video.embedder import Embedder
    from offmark_b200.video.extractor import Extractor
    from offmark_b200.video.memory_io import ArrayReader, ArrayWriter, BatchReader, BatchWriter
    from oracle import synth
    frames = np.stack([synth.random_bgr(128, 192, s) for s in range(14)])     # 5 batches of 3: more than the lanes in flight

    def encoder(payload=PAYLOAD):
        enc = DwtDctSvdEncoder()
        enc.read_wm(Shuffler(key=KEY).generate_wm(payload, enc.wm_capacity(frames[0].shape)))
        return enc
    ref_writer = ArrayWriter()
    Embedder(ArrayReader(list(frames)), encoder(), ref_writer).start()                      # per-frame loop
    pinned = torch.from_numpy(frames).pin_memory()
    for reader, writer in ((BatchReader(pinned.numpy()), BatchWriter(len(frames), frames[0].shape)),
                           (BatchReader(frames), BatchWriter(len(frames), frames[0].shape, pinned=False)),
                           (BatchReader(frames), ArrayWriter()), (ArrayReader(list(frames)), BatchWriter(len(frames), frames[0].shape))):
        Embedder(reader, encoder(), writer, batch_frames=3).start()
        assert len(writer.frames) == len(frames)
        for a, b in zip(ref_writer.frames, writer.frames):
            assert np.array_equal(a, b)
    other = 1 - PAYLOAD                             # second half of the clip carries another payload: order is visible
    other_writer = ArrayWriter()
    Embedder(ArrayReader(list(frames[5:])), encoder(other), other_writer).start()
    marked = np.stack(ref_writer.frames[:5] + other_writer.frames)
    want = Extractor(ArrayReader(list(marked)), DwtDctSvdDecoder(), DeShuffler(key=KEY).set_shape((8,)))
    want.start()
    for reader in (BatchReader(marked), BatchReader(torch.from_numpy(marked).pin_memory().numpy())):
        got = Extractor(reader, DwtDctSvdDecoder(), DeShuffler(key=KEY).set_shape((8,)), batch_frames=4)
        got.start()
        assert len(want.patterns) == len(got.patterns) == len(frames)
        for k, (a, b) in enumerate(zip(want.patterns, got.patterns)):
            assert np.array_equal(a, b) and np.array_equal(a, PAYLOAD if k < 5 else other)
    r = BatchReader(frames)                         # the reference's own protocol on the same object
    assert np.array_equal(r.read(), frames[0]) and np.array_equal(r.read_batch(100), frames[1:]) and r.read() is None
