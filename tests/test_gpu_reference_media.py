"""GPU: the reference's own media fixtures at FULL size (frame63.jpeg at 1080p, all 209 frames of in.mp4) through the
CUDA path, against what the reference itself produced (tests/golden/frame63_full.npz, in_mp4_all.npz: hashes, packed
bits and patterns written by oracle/make_golden.py, which ran the reference's classes) and against the oracle's marked
frames, which are first proven to still be the reference's output on this box (hash equality)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import bracket, dct8 as o_dct, dwt_dct_svd as o_svd, payload as o_pay
from parity import PAYLOAD, KEY, knife_edge_blocks, tile_mask_to_pixels
from test_oracle_golden import load_frame63, load_in_mp4

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bits(packed, n):
    return np.unpackbits(packed)[:n]


def test_frame63_whole_frame_dwtsvd(golden_dir):
    """The plugin flow of tests/mark.py / detect.py on the reference's 1080p still: fused rgb24 mark within 1 LSB of the
    reference's marked frame, float32 plugin encode within 2e-3, raw bits of clean / marked frames equal to the
    reference's arrays off the knife-edge blocks, patterns equal."""
    from b200wm import ops
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    img, g = load_frame63(golden_dir)
    h, w, _ = img.shape
    n = int(g["dwtsvd_nbits"])
    wm = o_pay.generate_wm(PAYLOAD, (1, n), KEY)
    yuv0 = bracket.to_yuv(img)
    want_f32 = o_svd.encode(yuv0.copy(), wm)
    want_u8 = bracket.from_yuv(want_f32.copy())
    assert _sha(want_u8) == str(g["dwtsvd_marked_u8_sha256"])            # the oracle IS the reference, on this box too
    _, edge_floor, _ = knife_edge_blocks(yuv0[:, :, 1])
    ok = ~tile_mask_to_pixels(edge_floor, (h, w))
    # fused rgb24 kernel (what Embedder.mark_frame runs)
    t = torch.from_numpy(img.copy()).to(DEV)
    packed, nb = ops.pack_bits(wm[0], device=DEV)
    ops.dwtsvd_embed_rgb8_(t, packed, nb)
    d = np.abs(t.cpu().numpy().astype(np.int16) - want_u8).max(axis=2)
    assert d[ok].max() <= 1 and (d > 0).mean() < 2e-3, (d[ok].max(), (d > 0).mean())       # tolerance: 1 LSB
    # float32 plugin object, in-place semantics
    enc = DwtDctSvdEncoder()
    enc.read_wm(wm)
    yuv = yuv0.copy()
    assert enc.encode(yuv) is yuv
    assert np.abs(yuv[:, :, 1] - want_f32[:, :, 1])[ok].max() < 2e-3
    # extraction against the reference's own bit arrays
    deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
    for name, src in (("clean", yuv0), ("marked_f32", want_f32), ("marked_u8", bracket.to_yuv(want_u8))):
        bits = DwtDctSvdDecoder().decode(src.copy())
        gold = _bits(g[f"dwtsvd_bits_{name}"], n)
        diff = np.flatnonzero(bits[0].astype(np.uint8) != gold)
        edge, _, _ = knife_edge_blocks(src[:, :, 1])
        assert all(c < edge.size and edge[c] for c in diff), (name, len(diff))
        assert np.array_equal(deg.degenerate(bits), g[f"dwtsvd_pattern_{name}"]), name
    raw, _ = ops.dwtsvd_extract_rgb8(torch.from_numpy(want_u8).to(DEV))
    diff = np.flatnonzero(ops.unpack_bits(raw, n)[0] != _bits(g["dwtsvd_bits_marked_u8"], n))
    edge, _, _ = knife_edge_blocks(bracket.to_yuv(want_u8)[:, :, 1])
    assert all(edge[c] for c in diff)


def test_frame63_whole_frame_dct8(golden_dir):
    """DctEncoder / DctDecoder on the same still: every differing block is a named float32 tie (oracle/dct8.py:tie_blocks),
    bits against the reference's arrays, identical patterns."""
    from offmark_b200.embed.dct_encoder import DctEncoder
    from offmark_b200.extract.dct_decoder import DctDecoder
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    img, g = load_frame63(golden_dir)
    h, w, _ = img.shape
    by, bx = h // 8, w // 8
    n = int(g["dct8_nbits"])
    wm = o_pay.generate_wm(PAYLOAD, (1, n), KEY)
    yuv0 = bracket.to_yuv(img)
    want = o_dct.encode(yuv0.copy(), wm)
    assert _sha(want[:, :, 1]) == str(g["dct8_marked_f32_ch1_sha256"])
    enc = DctEncoder()
    enc.read_wm(wm)
    got = enc.encode(yuv0.copy())
    ties = o_dct.tie_blocks(yuv0)
    err = np.abs(got[:, :, 1] - want[:, :, 1]).reshape(by, 8, bx, 8).max(axis=(1, 3))
    bad = (err >= 2e-3) & ~(ties["mask"] | ties["sign"] | ties["floor"])
    assert not bad.any(), f"{int(bad.sum())} blocks differ without a float32 tie"
    assert (err >= 2e-3).mean() < 6e-2
    deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
    marked_u8 = bracket.from_yuv(want.copy())
    for name, src in (("clean", yuv0), ("marked_f32", want), ("marked_u8", bracket.to_yuv(marked_u8))):
        bits = DctDecoder().decode(src.copy())
        t2 = o_dct.tie_blocks(src)
        differ = (bits[0].astype(np.uint8) != _bits(g[f"dct8_bits_{name}"], n)).reshape(by, bx)
        assert not (differ & ~(t2["mask"] | t2["round"])).any(), name
        assert differ.mean() < 1e-2
        assert np.array_equal(deg.degenerate(bits), g[f"dct8_pattern_{name}"]), name


def test_in_mp4_all_209_frames(golden_dir):
    """mark.py + detect.py on every frame of the reference's clip in ONE batch: fused rgb24 marks within 1 LSB of the
    reference's marked frames (the oracle's, hash-checked), raw bits of the reference-marked frames against the
    reference decoder's, the 209 per-frame patterns identical, and the clip vote."""
    from b200wm import ops
    from b200wm.vote import SegmentVote
    frames, g = load_in_mp4(golden_dir)
    n_frames, (h, w, _) = len(frames), frames[0].shape
    n = int(g["nbits"])
    wm = o_pay.generate_wm(PAYLOAD, (1, n), KEY)
    ref_marked = []
    for i, f in enumerate(frames):
        m = bracket.mark_frame(f, lambda y: o_svd.encode(y, wm))
        assert _sha(m) == str(g["marked_sha256"][i]), i
        ref_marked.append(m)
    ref_marked = np.stack(ref_marked)
    t = torch.from_numpy(np.stack(frames)).to(DEV)
    packed, nb = ops.pack_bits(wm[0], device=DEV)
    ops.dwtsvd_embed_rgb8_(t, packed, nb)
    got = t.cpu().numpy()
    worst, frac = 0, 0.0
    for i in range(n_frames):
        _, edge_floor, _ = knife_edge_blocks(bracket.to_yuv(frames[i])[:, :, 1], check=False)
        assert edge_floor.sum() <= 2
        ok = ~tile_mask_to_pixels(edge_floor, (h, w))
        d = np.abs(got[i].astype(np.int16) - ref_marked[i]).max(axis=2)
        worst, frac = max(worst, int(d[ok].max())), frac + float((d > 0).mean()) / n_frames
    assert worst <= 1 and frac < 2e-3, (worst, frac)                       # tolerance: 1 LSB
    raw, counts = ops.dwtsvd_extract_rgb8(torch.from_numpy(ref_marked).to(DEV), payload_len=8)
    bits = ops.unpack_bits(raw, n)
    mismatches = 0
    for i in range(n_frames):
        diff = np.flatnonzero(bits[i] != _bits(g["bits_marked"][i], n))
        if diff.size:
            edge, _, _ = knife_edge_blocks(bracket.to_yuv(ref_marked[i])[:, :, 1], check=False)
            assert all(edge[c] for c in diff), i
            mismatches += diff.size
    assert mismatches <= 4
    perm = torch.from_numpy(o_pay.permutation(8, KEY).astype(np.int32)).to(DEV)
    patterns, words = ops.vote_finish(counts, n, perm)
    assert np.array_equal(patterns.cpu().numpy(), g["patterns"])
    pattern, freq, _, seen = SegmentVote(1, 8, DEV).add(words).result()[0]
    want_p, want_f = o_pay.pattern_vote(list(g["patterns"]))
    assert np.array_equal(pattern, want_p) and freq == want_f and seen == n_frames
