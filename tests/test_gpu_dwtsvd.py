"""GPU: Haar-DWT / block-SVD embed and extract through the C ABI vs the oracle and the golden fixtures."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import bracket, dwt_dct_svd as o_svd, payload as o_pay, synth
from parity import PAYLOAD, KEY, knife_edge_blocks, tile_mask_to_pixels, flat_tiles

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1], ids=["tma", "ldg"])
def kernel_path(request):
    """Every test runs on both kernel families: TMA-staged persistent (auto) and vectorised loads."""
    from b200wm import ops
    ops.set_path(request.param)
    yield request.param
    ops.set_path(0)


def _dev():
    return torch.device("cuda:0")


def _wm(shape_hw):
    h, w = shape_hw
    return o_pay.generate_wm(PAYLOAD, (1, h * w // 64), KEY)


def _extract_bits(plane_t, n):
    from b200wm import ops
    raw, _ = ops.dwtsvd_extract(plane_t)
    torch.cuda.synchronize()
    return ops.unpack_bits(raw, n)


def _assert_bits_match(got, want, plane_f32, what):
    """Bit-exact, except on blocks whose sigma_0 sits on a quantisation boundary (parity.py)."""
    want = np.asarray(want).reshape(-1).astype(np.uint8)
    got = np.asarray(got).reshape(-1).astype(np.uint8)
    assert got.shape == want.shape
    diff = np.flatnonzero(got != want)
    if diff.size:
        edge, _, s64 = knife_edge_blocks(plane_f32)
        off_edge = [c for c in diff if c >= edge.size or not edge[c]]
        assert not off_edge, f"{what}: {len(off_edge)} raw bits differ away from quantisation boundaries, e.g. block {off_edge[:5]} sigma {s64[off_edge[:5]]}"
    return diff.size


@pytest.mark.parametrize("h,w", [(1080, 1920), (240, 320), (64, 64), (8, 8), (40, 72), (37, 53), (100, 132), (7, 9), (12, 20)])
def test_extract_u8_planes_vs_oracle(h, w):
    from b200wm import ops
    plane = synth.luma_plane_u8(h, w, 3, 99)
    want = o_svd.extract_plane(plane)
    got = _extract_bits(torch.from_numpy(plane).to(_dev()), want.size)
    _assert_bits_match(got, want, plane.astype(np.float32), f"u8 {h}x{w}")
    block_num, tiles, words = ops.geometry(h, w)
    assert got.shape == (1, block_num)
    assert not got[0, tiles:].any()          # surplus bits are zero like decoder.py:14-15


def test_extract_unaligned_and_strided_views_take_generic_path():
    """Same bits whether the plane is 8-byte aligned (vector path) or not (generic path)."""
    plane = synth.luma_plane_u8(120, 200, 1, 5)
    want = o_svd.extract_plane(plane)
    big = torch.zeros((130, 211), dtype=torch.uint8, device=_dev())
    big[3:123, 5:205] = torch.from_numpy(plane).to(_dev())
    view = big[3:123, 5:205]                  # pitch 211, base offset 5: unaligned
    assert view.data_ptr() % 8 != 0
    got = _extract_bits(view, want.size)
    _assert_bits_match(got, want, plane.astype(np.float32), "unaligned view")


def test_extract_full_range_content_incl_flat_and_zero_blocks():
    plane = synth.full_range_plane_u8(256, 384, 11)
    want = o_svd.extract_plane(plane)
    got = _extract_bits(torch.from_numpy(plane).to(_dev()), want.size)
    _assert_bits_match(got, want, plane.astype(np.float32), "full range")
    # The flat 255 quarter: exact sums give sigma_0 = 2040 = 136*15, a boundary; the reference's float32 Haar band
    # gives 2039.99988 -> bit 1.  The kernels follow the reference on flat tiles (dwtsvd_tile.cuh), unmasked:
    flat = flat_tiles(plane)
    assert flat.sum() > 100 and np.array_equal(got[0, :flat.size][flat], want[0, :flat.size][flat].astype(np.uint8))
    white = flat & (plane[::8, ::8][:plane.shape[0] // 8, :plane.shape[1] // 8].reshape(-1) == 255)
    assert white.any() and (got[0, :flat.size][white] == 1).all()


def test_sigma_matches_float64_truth():
    from b200wm import ops
    for plane in (synth.luma_plane_u8(240, 320, 0, 1), synth.full_range_plane_u8(128, 128, 2)):
        yuv = np.zeros(plane.shape + (3,), dtype=np.float32)
        yuv[:, :, 1] = plane
        _, s64 = o_svd.decode_sigma(yuv)
        sig = ops.dwtsvd_sigma(torch.from_numpy(plane).to(_dev()))[0].cpu().numpy()
        np.testing.assert_allclose(sig, s64, rtol=3e-7, atol=1e-6)


def test_block_dct_is_dropped_without_changing_sigma():
    """The production kernels skip the reference's 4x4 block DCT (svd(cv2.dct(block)),
    dwt_dct_svd_decoder.py:35) because it is orthogonal.  b200wm_dwtsvd_sigma_dct keeps it (4-point
    butterflies): both agree with each other and with the float64 truth of the reference's expression."""
    from b200wm import ops
    for plane in (synth.luma_plane_u8(240, 320, 0, 1), synth.full_range_plane_u8(128, 128, 2)):
        yuv = np.zeros(plane.shape + (3,), dtype=np.float32)
        yuv[:, :, 1] = plane
        _, s64 = o_svd.decode_sigma(yuv)
        t = torch.from_numpy(plane).to(_dev())
        fast = ops.dwtsvd_sigma(t)[0].cpu().numpy()
        faithful = ops.dwtsvd_sigma_dct(t)[0].cpu().numpy()
        np.testing.assert_allclose(faithful, s64, rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(faithful, fast, rtol=1e-6, atol=1e-5)
        tf = torch.from_numpy(yuv).to(_dev())
        np.testing.assert_allclose(ops.dwtsvd_sigma_dct(tf, channel=1)[0].cpu().numpy(), s64, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("h,w", [(1080, 1920), (240, 320), (37, 53), (100, 132), (8, 8)])
def test_embed_u8_planes_within_1_lsb_of_oracle(h, w):
    from b200wm import ops
    plane = synth.luma_plane_u8(h, w, 7, 2024)
    wm = _wm((h, w))
    want = o_svd.embed_plane_u8(plane, wm[0])
    t = torch.from_numpy(plane).to(_dev())
    packed, n = ops.pack_bits(wm[0], device=_dev())
    ops.dwtsvd_embed_(t, packed, n)
    got = t.cpu().numpy()
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    _, edge_floor, _ = knife_edge_blocks(plane.astype(np.float32))
    ok_px = ~tile_mask_to_pixels(edge_floor, plane.shape)
    assert diff[ok_px].max(initial=0) <= 1, "marked plane differs from the reference embedder by more than 1 LSB"
    assert (diff[ok_px] > 0).mean() < 1e-3
    nr, nc = o_svd.block_grid(h, w)
    assert np.array_equal(got[nr * 8:], plane[nr * 8:]) and np.array_equal(got[:, nc * 8:], plane[:, nc * 8:])
    # cross checks: each extractor reads the other's embedding
    tiles = nr * nc
    ref_reads_ours = o_svd.extract_plane(got)[0]
    we_read_ref = _extract_bits(torch.from_numpy(want).to(_dev()), ref_reads_ours.size)[0]
    ref_reads_ref = o_svd.extract_plane(want)[0]
    assert (ref_reads_ours[:tiles] == wm[0][:tiles]).mean() >= (ref_reads_ref[:tiles] == wm[0][:tiles]).mean() - 0.01
    _assert_bits_match(we_read_ref, ref_reads_ref, want.astype(np.float32), "gpu extract of reference embed")
    for bits in (ref_reads_ours, we_read_ref):
        assert np.array_equal(o_pay.degenerate(bits.reshape(1, -1), 8, KEY), PAYLOAD) or tiles < 64


def test_embed_full_range_clips_like_reference():
    from b200wm import ops
    plane = synth.full_range_plane_u8(256, 384, 4)
    wm = _wm(plane.shape)
    want = o_svd.embed_plane_u8(plane, wm[0])
    t = torch.from_numpy(plane).to(_dev())
    packed, n = ops.pack_bits(wm[0], device=_dev())
    ops.dwtsvd_embed_(t, packed, n)
    got = t.cpu().numpy()
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    _, edge_floor, s64 = knife_edge_blocks(plane.astype(np.float32))
    # blocks whose two largest singular values nearly coincide have no unique singular pair
    yuv = np.zeros(plane.shape + (3,), dtype=np.float32); yuv[:, :, 1] = plane
    ca, _ = __import__("oracle.haar", fromlist=["x"]).dwt2_haar(yuv[:, :, 1])
    blocks, _, _ = o_svd._to_blocks(ca, 4)
    s = np.linalg.svd(blocks.astype(np.float64), compute_uv=False)
    ambiguous = ((s[:, 0] - s[:, 1]) < 1e-3 * s[:, 0]) & (s[:, 0] > 0)      # all-zero blocks are NOT masked: svd(0) is exact
    ok_px = ~tile_mask_to_pixels(edge_floor | ambiguous, plane.shape)
    assert diff[ok_px].max(initial=0) <= 1
    assert (edge_floor | ambiguous).mean() < 2e-3, (edge_floor.mean(), ambiguous.mean())
    # all-zero blocks: svd(0) = (I, 0, I) puts the mark on the DC term -> +0 / +1 for bit 0 / 1, exactly like the
    # reference; flat 255 blocks: the reference's floor is 135 (sigma_0 = 2039.99988), -1 / 0 for bit 0 / 1 - unmasked
    flat = tile_mask_to_pixels(flat_tiles(plane), plane.shape)
    assert flat.sum() > 64 * 100 and np.array_equal(got[flat], want[flat])
    zero_px = tile_mask_to_pixels(s[:, 0] == 0, plane.shape)
    assert zero_px.any() and set(np.unique(got[zero_px])) == {0, 1}


def test_embed_out_of_place_and_per_frame_rows():
    from b200wm import ops
    n_frames, h, w = 6, 64, 96
    planes = np.stack([synth.luma_plane_u8(h, w, f, 77) for f in range(n_frames)])
    payloads = [o_pay.payload_for_segment(s) for s in (5, 200, 77)]
    rows = np.stack([o_pay.generate_wm(p, (1, h * w // 64), KEY)[0] for p in payloads])
    frame_row = np.array([0, 0, 1, 1, 2, 2], dtype=np.int32)
    src = torch.from_numpy(planes).to(_dev())
    dst = src.clone()
    packed, n = ops.pack_bits(rows, device=_dev())
    ops.dwtsvd_embed_(src, packed, n, frame_wm_row=torch.from_numpy(frame_row).to(_dev()), out=dst)
    assert np.array_equal(src.cpu().numpy(), planes)           # source untouched
    got = dst.cpu().numpy()
    for f in range(n_frames):
        want = o_svd.embed_plane_u8(planes[f], rows[frame_row[f]])
        assert np.abs(got[f].astype(np.int16) - want).max() <= 1
    raw, counts = ops.dwtsvd_extract(dst, payload_len=8)
    bits = ops.unpack_bits(raw, h * w // 64)
    for f in range(n_frames):
        assert np.array_equal(o_pay.degenerate(bits[f].reshape(1, -1), 8, KEY), payloads[frame_row[f] if True else 0])
        assert counts[f].cpu().numpy().tolist() == [int(bits[f][i::8].sum()) for i in range(8)]


def test_short_watermark_raises_index_error_like_reference():
    from b200wm import ops
    t = torch.zeros((64, 64), dtype=torch.uint8, device=_dev())
    packed, n = ops.pack_bits(np.zeros(10, dtype=np.int64), device=_dev())
    with pytest.raises(IndexError):
        ops.dwtsvd_embed_(t, packed, n)


def test_float_interleaved_frames_vs_oracle(golden_dir):
    """The reference layout: float32 H x W x 3, channel 1 marked (golden frame63 crop)."""
    from b200wm import ops
    g = np.load(os.path.join(golden_dir, "frame63_crop.npz"))
    frame = g["bgr"]
    yuv0 = bracket.to_yuv(frame)
    wm = o_pay.generate_wm(PAYLOAD, o_svd.wm_capacity(frame.shape), KEY)
    want = o_svd.encode(yuv0.copy(), wm)
    t = torch.from_numpy(yuv0).to(_dev())
    packed, n = ops.pack_bits(wm[0], device=_dev())
    ops.dwtsvd_embed_(t, packed, n, channel=1)
    got = t.cpu().numpy()
    assert np.array_equal(got[:, :, 0], yuv0[:, :, 0]) and np.array_equal(got[:, :, 2], yuv0[:, :, 2])
    _, edge_floor, _ = knife_edge_blocks(yuv0[:, :, 1])
    ok = ~tile_mask_to_pixels(edge_floor, yuv0.shape[:2])
    err = np.abs(got[:, :, 1] - want[:, :, 1])
    assert err[ok].max() < 2e-3, err[ok].max()
    h, w = g["dwtsvd_marked_f32_ch1_window"].shape
    assert np.abs(got[:h, :w, 1] - g["dwtsvd_marked_f32_ch1_window"])[ok[:h, :w]].max() < 2e-3
    # after the uint8 bracket the frames agree within 1 LSB
    ours_u8, ref_u8 = bracket.from_yuv(got.copy()), bracket.from_yuv(want.copy())
    d = np.abs(ours_u8.astype(np.int16) - ref_u8)
    assert d[ok].max() <= 1 and (d > 0).mean() < 1e-3
    # extraction from the three golden sources
    nbits = int(g["dwtsvd_nbits"])
    for name, src in (("clean", yuv0), ("marked_f32", want), ("marked_u8", bracket.to_yuv(ref_u8))):
        raw, _ = ops.dwtsvd_extract(torch.from_numpy(src).to(_dev()), channel=1)
        bits = ops.unpack_bits(raw, nbits)
        gold = np.unpackbits(g[f"dwtsvd_bits_{name}"])[:nbits]
        _assert_bits_match(bits, gold, src[:, :, 1], f"golden {name}")


def test_golden_u8_plane_1080p(golden_dir):
    from b200wm import ops
    g = np.load(os.path.join(golden_dir, "u8plane_1080p.npz"))
    y = synth.luma_plane_u8(1080, 1920, int(g["frame_index"]), int(g["seed"]))
    n = 32400
    t = torch.from_numpy(y).to(_dev())
    _assert_bits_match(_extract_bits(t, n), np.unpackbits(g["bits_clean"])[:n], y.astype(np.float32), "golden clean")
    wm = _wm((1080, 1920))
    packed, nb = ops.pack_bits(wm[0], device=_dev())
    ops.dwtsvd_embed_(t, packed, nb)
    got = t.cpu().numpy()
    d = got[:64].astype(np.int16) - y[:64]
    assert np.abs(d - g["marked_minus_src_rows_0_64"]).max() <= 1
    same = hashlib.sha256(got.tobytes()).hexdigest() == str(g["marked_sha256"])
    print("marked plane identical to the reference's byte for byte:", same)
    bits = _extract_bits(t, n)
    assert np.array_equal(o_pay.degenerate(bits.reshape(1, -1), 8, KEY), g["pattern_marked"])


@pytest.mark.parametrize("h,w,n_frames", [(2160, 3840, 3), (1080, 1920, 8)])
def test_full_size_properties(h, w, n_frames):
    """BASELINE sizes: embed -> extract round trip, idempotence of the quantiser, untouched planes."""
    from b200wm import ops
    gen = torch.Generator(device=_dev()).manual_seed(5)
    planes = torch.randint(16, 236, (n_frames, h, w), dtype=torch.uint8, device=_dev(), generator=gen)
    planes = (planes.float() * 0.25 + 96 + 40 * torch.sin(torch.arange(w, device=_dev()) / 37.0)).round().clamp(16, 235).to(torch.uint8)
    wm = _wm((h, w))
    packed, nb = ops.pack_bits(wm[0], device=_dev())
    marked = planes.clone()
    ops.dwtsvd_embed_(marked, packed, nb)
    assert (marked.int() - planes.int()).abs().max().item() <= 8      # |delta sigma| < 15 -> |delta pixel| < 7.5
    raw, counts = ops.dwtsvd_extract(marked, payload_len=8)
    bits = ops.unpack_bits(raw, h * w // 64)
    acc = (bits == wm[0][None, :]).mean(axis=1)
    assert acc.min() > 0.9
    perm = torch.from_numpy(o_pay.permutation(8, KEY).astype(np.int32)).to(_dev())
    patterns, packed_p = ops.vote_finish(counts, h * w // 64, perm)
    assert (patterns.cpu().numpy() == PAYLOAD[None, :]).all()
    assert packed_p.cpu().tolist() == [0b01100101] * n_frames
    # embedding the same mark again must move (almost) nothing: sigma_0 already sits on the lattice
    again = marked.clone()
    ops.dwtsvd_embed_(again, packed, nb)
    assert (again.int() - marked.int()).abs().float().mean().item() < 0.6
    # checksum of counts equals the popcount of the raw bits
    assert counts.sum(dim=1).cpu().tolist() == [int(b.sum()) for b in bits]


def test_no_out_of_bounds_writes_guard_bytes():
    """compute-sanitizer is closed on this pool, so bounds are checked with canaries: the planes sit
    inside a larger buffer whose guard bytes (before, after, and the chroma planes between the luma
    planes of an I420 layout) must come back untouched from embed and extract, on both kernel paths."""
    from b200wm import ops
    n, h, w = 5, 72, 1536                       # 192 tiles per row: TMA-eligible, 9 tile rows
    frame_bytes = h * w * 3 // 2
    guard = 4096
    buf = torch.full((guard + n * frame_bytes + guard,), 0xA5, dtype=torch.uint8, device=_dev())
    frames = buf[guard:guard + n * frame_bytes].view(n, frame_bytes)
    gen = torch.Generator(device=_dev()).manual_seed(1)
    y = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=_dev(), generator=gen)
    planes = frames.as_strided((n, h, w), (frame_bytes, w, 1))
    planes.copy_(y)
    wm = _wm((h, w))
    packed, nb = ops.pack_bits(wm[0], device=_dev())
    ops.dwtsvd_embed_(planes, packed, nb)
    raw, counts = ops.dwtsvd_extract(planes, payload_len=8)
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    assert (host[:guard] == 0xA5).all() and (host[-guard:] == 0xA5).all()
    chroma = host[guard:guard + n * frame_bytes].reshape(n, frame_bytes)[:, h * w:]
    assert (chroma == 0xA5).all(), "embed wrote outside the luma planes"
    marked = planes.cpu().numpy()
    for f in range(n):
        want = o_svd.embed_plane_u8(y[f].cpu().numpy(), wm[0])
        _, edge_floor, _ = knife_edge_blocks(y[f].cpu().numpy().astype(np.float32))
        ok = ~tile_mask_to_pixels(edge_floor, (h, w))
        assert np.abs(marked[f].astype(np.int16) - want)[ok].max() <= 1
    # the output buffers are exactly sized too
    assert raw.shape == (n, (h * w // 64 + 31) // 32) and counts.shape == (n, 8)


def test_random_geometries_both_paths_agree():
    """Random plane sizes, pitches, batch sizes and per-frame watermark rows: the TMA path (whole
    strips, row-wise copies of pitched planes, column chunks of wide planes with an uneven last
    chunk), the vectorised path and the generic path (unaligned view) must give byte-identical
    planes and bits."""
    from b200wm import ops
    rng = np.random.RandomState(12345)
    widths = [64, 320, 528, 784, 1000, 1024, 1040, 1080, 1936, 2048, 2064, 3840, 4112, 6160]   # <= 128 tiles: two tile rows per item; 1000 / 1080: odd tile counts (125, 135), rows only 8-byte aligned; 258, 480, 514, 770 tiles: 2, 2, 3, 4 chunks
    for case in range(28):
        n = int(rng.randint(1, 6))
        h = int(rng.choice([8, 16, 24, 40, 64, 72, 136, 270]))
        w = widths[case % len(widths)]
        base = rng.randint(0, 256, (n, h, w)).astype(np.uint8)
        rows = rng.randint(0, 2, (3, max(1, h * w // 64)))
        frame_row = rng.randint(0, 3, n).astype(np.int32)
        outs = []
        for mode in ("auto", "pitched16", "ldg", "unaligned"):
            ops.set_path(1 if mode == "ldg" else 0)
            if mode == "unaligned":
                big = torch.zeros((n, h + 2, w + 13), dtype=torch.uint8, device=_dev())
                t = big[:, 1:h + 1, 5:w + 5]
                t.copy_(torch.from_numpy(base))
            elif mode == "pitched16":       # rows 16-byte aligned but not tight: TMA path, one bulk copy per sample row
                big = torch.full((n, h + 3, w + 48), 0x5A, dtype=torch.uint8, device=_dev())
                t = big[:, 2:h + 2, 16:w + 16]
                t.copy_(torch.from_numpy(base))
            else:
                t = torch.from_numpy(base.copy()).to(_dev())
            packed, nb = ops.pack_bits(rows, device=_dev())
            ops.dwtsvd_embed_(t, packed, nb, frame_wm_row=torch.from_numpy(frame_row).to(_dev()))
            raw, counts = ops.dwtsvd_extract(t, payload_len=8)
            outs.append((t.cpu().numpy(), raw.cpu().numpy(), counts.cpu().numpy()))
            if mode == "pitched16":         # nothing outside the view was written
                big_h = big.cpu().numpy()
                big_h[:, 2:h + 2, 16:w + 16] = 0x5A
                assert (big_h == 0x5A).all(), (case, n, h, w)
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert np.array_equal(a, b), (case, n, h, w)
    ops.set_path(0)


@pytest.mark.parametrize("h,w,pad", [(1080, 1920, 0), (240, 320, 0), (70, 90, 3), (37, 53, 0)])
def test_multi_copy_embed_equals_single_embeds(h, w, pad):
    """b200wm_dwtsvd_embed_copies: one read, N marked copies (tests/mark_video_to_hls.py:330-354), each
    bit-identical to a single embed of its payload row - with per-(frame, copy) row tables, pitched
    planes (generic path) and tile-uncovered edges - and every copy's payload decodes."""
    from b200wm import ops
    DEV = _dev()
    n_frames, n_copies = 3, 4
    rng = np.random.RandomState(h + w)
    frames = rng.randint(0, 256, (n_frames, h, w + pad)).astype(np.uint8)
    frames[0, :16, :32] = 0           # all-zero blocks
    frames[1, :8, :8] = 255
    src = torch.from_numpy(frames).to(DEV)[:, :, :w]
    n = h * w // 64
    payloads = [o_pay.payload_for_segment_copy(s, c) for s in range(2) for c in range(n_copies)]
    wm = np.stack([o_pay.generate_wm(p, (1, n), KEY)[0] for p in payloads])
    packed, ln = ops.pack_bits(wm, device=DEV)
    # frame f belongs to segment f % 2: row = segment * n_copies + copy
    table = torch.tensor([[(f % 2) * n_copies + c for c in range(n_copies)] for f in range(n_frames)], dtype=torch.int32, device=DEV)
    out = ops.dwtsvd_embed_copies(src, packed, ln, n_copies, copy_wm_row=table)
    assert out.shape == (n_copies, n_frames, h, w)
    for c in range(n_copies):
        single = src.clone()
        ops.dwtsvd_embed_(single, packed, ln, frame_wm_row=table[:, c].contiguous())
        assert torch.equal(out[c], single), f"copy {c} differs from a single embed"
        raw, counts = ops.dwtsvd_extract(out[c], payload_len=8)
        perm = torch.from_numpy(o_pay.permutation(8, KEY).astype(np.int32)).to(DEV)
        patterns, _ = ops.vote_finish(counts, n, perm)
        if h >= 240:
            for f in range(n_frames):
                assert np.array_equal(patterns[f].cpu().numpy(), payloads[(f % 2) * n_copies + c])
    # default table: row c for every frame
    out2 = ops.dwtsvd_embed_copies(src, packed, ln, 2)
    for c in range(2):
        single = src.clone()
        ops.dwtsvd_embed_(single, packed, ln, frame_wm_row=torch.full((n_frames,), c, dtype=torch.int32, device=DEV))
        assert torch.equal(out2[c], single)
    if n > 32:
        with pytest.raises(IndexError):          # watermark shorter than the block count, as embed/dwt_dct_svd_encoder.py:36
            ops.dwtsvd_embed_copies(src, packed[:, :1].contiguous(), 32, 2)


@pytest.mark.parametrize("which", ["y", "u", "v"])
def test_i420_native_planes(which):
    """mark / read a plane of planar yuv420p frames in place (the wire format of
    video/frame_writer.py:34) through strided views: the chosen plane matches the oracle within 1 LSB,
    its payload decodes, and the two other planes are not touched."""
    from b200wm import ops
    DEV = _dev()
    n, h, w = 4, 1080, 1920
    fb = h * w * 3 // 2
    rng = np.random.RandomState(7)
    host = np.empty((n, fb), dtype=np.uint8)
    for f in range(n):
        host[f, :h * w] = synth.luma_plane_u8(h, w, f, 11).reshape(-1)
        host[f, h * w:] = np.clip(np.rint(128 + 20 * np.sin(np.arange(h * w // 2) / 977.0) + rng.normal(0, 3, h * w // 2)), 16, 240)
    frames = torch.from_numpy(host).to(DEV)
    plane = ops.i420_plane(frames, h, w, which)
    ph, pw = plane.shape[1], plane.shape[2]
    wm = o_pay.generate_wm(PAYLOAD, (1, ph * pw // 64), KEY)
    packed, nb = ops.pack_bits(wm[0], device=DEV)
    ops.dwtsvd_embed_(plane, packed, nb)
    raw, counts = ops.dwtsvd_extract(plane, payload_len=8)
    perm = torch.from_numpy(o_pay.permutation(8, KEY).astype(np.int32)).to(DEV)
    patterns, _ = ops.vote_finish(counts, ph * pw // 64, perm)
    assert (patterns.cpu().numpy() == PAYLOAD[None, :]).all()
    out = frames.cpu().numpy()
    lo = {"y": 0, "u": h * w, "v": h * w + ph * pw}[which]
    untouched = np.ones(fb, bool)
    untouched[lo:lo + ph * pw] = False
    assert np.array_equal(out[:, untouched], host[:, untouched]), "another plane was written"
    for f in (0, n - 1):
        src_plane = host[f, lo:lo + ph * pw].reshape(ph, pw)
        want = o_svd.embed_plane_u8(src_plane, wm[0])
        _, edge_floor, _ = knife_edge_blocks(src_plane.astype(np.float32))
        ok = ~tile_mask_to_pixels(edge_floor, (ph, pw))
        got = out[f, lo:lo + ph * pw].reshape(ph, pw)
        assert np.abs(got.astype(np.int16) - want)[ok].max() <= 1        # tolerance: 1 LSB (north_star)
        _assert_bits_match(ops.unpack_bits(raw[f:f + 1], ph * pw // 64), o_svd.extract_plane(got), got.astype(np.float32), f"i420 {which}")


def test_new_entry_points_reject_bad_arguments():
    """Status codes of the copies / resize entry points map to the exceptions of b200wm._lib.check."""
    from b200wm import ops
    DEV = _dev()
    planes = torch.zeros((2, 64, 512), dtype=torch.uint8, device=DEV)
    packed, n = ops.pack_bits(np.zeros((2, 64 * 512 // 64), dtype=np.int64), device=DEV)
    with pytest.raises(ValueError):                      # more copies than watermark rows and no row table
        ops.dwtsvd_embed_copies(planes, packed, n, 3)
    with pytest.raises(ValueError):                      # float planes: the multi-copy kernel is uint8 only
        ops.dwtsvd_embed_copies(planes.float(), packed, n, 2)
    with pytest.raises(ValueError):                      # row table of the wrong shape
        ops.dwtsvd_embed_copies(planes, packed, n, 2, copy_wm_row=torch.zeros((2, 3), dtype=torch.int32, device=DEV))
    assert ops.dwtsvd_embed_copies(planes, packed, n, 0).shape[0] == 0          # zero copies: nothing to do
    empty = ops.attack_resize(planes[:0], (256, 32), ops.INTER_AREA)            # zero frames: nothing to do
    assert empty.shape == (0, 32, 256)
    with pytest.raises(ValueError):
        ops.i420_plane(torch.zeros((2, 100), dtype=torch.uint8, device=DEV), 64, 512)


def _content_classes(h, w):
    """1080p planes of different character: what decides sigma_0 and the conditioning of its singular pair."""
    rng = np.random.RandomState(404)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    crop = np.load(os.path.join(os.path.dirname(__file__), "golden", "frame63_crop.npz"))["bgr"]
    tiled = np.tile(crop[:, :, 1], (-(-h // crop.shape[0]), -(-w // crop.shape[1])))[:h, :w]
    return {
        "natural image (reference fixture, tiled)": tiled.copy(),
        "smooth gradient": np.clip(16 + 200 * (xx / w) * (yy / h) + 8, 0, 255).astype(np.uint8),
        "white noise": rng.randint(0, 256, (h, w)).astype(np.uint8),
        "dark with noise": np.clip(rng.normal(6, 3, (h, w)), 0, 255).astype(np.uint8),
        "near saturation": np.clip(rng.normal(250, 4, (h, w)), 0, 255).astype(np.uint8),
        "checkerboard 4 px": (((xx // 4 + yy // 4) % 2) * 163 + 37).astype(np.uint8),     # sigma_0 = 948: off the 15-lattice
        "vertical edges": np.where((xx // 37) % 2 == 0, 40, 215).astype(np.uint8),
    }


@pytest.mark.parametrize("name", list(_content_classes(8, 8)))
def test_full_1080p_parity_across_content_classes(name):
    """Whole 1080p planes of several content classes against the oracle: marked plane within 1 LSB (outside the
    blocks whose floor quotient is decided inside the reference's own float32 rounding), raw bits of the marked
    plane identical except on quantisation-boundary blocks, voted payload identical."""
    from b200wm import ops
    h, w = 1080, 1920
    plane = _content_classes(h, w)[name]
    wm = _wm((h, w))
    want = o_svd.embed_plane_u8(plane, wm[0])
    t = torch.from_numpy(plane.copy()).to(_dev())
    packed, n = ops.pack_bits(wm[0], device=_dev())
    ops.dwtsvd_embed_(t, packed, n)
    got = t.cpu().numpy()
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    edge_bit, edge_floor, _ = knife_edge_blocks(plane.astype(np.float32))
    ok_px = ~tile_mask_to_pixels(edge_floor, plane.shape)
    assert diff[ok_px].max(initial=0) <= 1, name                      # tolerance: 1 LSB (north_star)
    # noise planes have ill-conditioned singular vectors (sigma_0 ~ sigma_1): more last-bit ties, still 1 LSB
    assert (diff[ok_px] > 0).mean() < (2e-2 if "noise" in name or "saturation" in name else 1e-3), (name, (diff[ok_px] > 0).mean())
    raw, counts = ops.dwtsvd_extract(t, payload_len=8)
    bits = ops.unpack_bits(raw, h * w // 64)
    want_bits = o_svd.extract_plane(got)
    _assert_bits_match(bits, want_bits, got.astype(np.float32), name)
    perm = torch.from_numpy(o_pay.permutation(8, KEY).astype(np.int32)).to(_dev())
    patterns, _ = ops.vote_finish(counts, h * w // 64, perm)
    assert np.array_equal(patterns[0].cpu().numpy(), o_pay.degenerate(want_bits, 8, KEY)), name


# ---- deterministic cases that no mask hides: all-zero and flat planes ---------------------------------------
FLAT_VALUES = (0, 1, 15, 16, 30, 45, 120, 128, 235, 240, 255)      # 15k/8-multiples sit ON a floor boundary (sigma_0 = 8v)


@pytest.mark.parametrize("h,w", [(64, 1920), (64, 2048), (32, 960), (16, 3840), (40, 72)])
def test_zero_and_flat_planes_equal_the_oracle_exactly(h, w):
    """All-zero and flat planes, payload bits 0 and 1: the reference is deterministic here (svd(0) = (I, 0, I);
    a constant block is exact in cv2.dct and LAPACK, sigma_0 = 16*fl(c*fl(c*v)) with the float32 Haar tap c),
    so marked planes and raw bits must EQUAL the oracle's - no tolerance, no mask - on the TMA kernels (whole
    strips, narrow planes, column chunks) and the vectorised-load kernels.  The marked planes are flat again
    (uniform increment), so extracting from them exercises the same rule a second time."""
    from b200wm import ops
    n = h * w // 64
    wm = _wm((h, w))[0]
    packed, nb = ops.pack_bits(np.stack([wm, np.zeros_like(wm), np.ones_like(wm)]), device=_dev())
    planes = np.stack([np.full((h, w), v, dtype=np.uint8) for v in FLAT_VALUES])
    src = torch.from_numpy(planes).to(_dev())
    clean_bits = _extract_bits(src, n)
    for row, bits_row in enumerate((wm, np.zeros_like(wm), np.ones_like(wm))):
        t = src.clone()
        ops.dwtsvd_embed_(t, packed, nb, frame_wm_row=torch.full((len(FLAT_VALUES),), row, dtype=torch.int32, device=_dev()))
        got = t.cpu().numpy()
        marked_bits = _extract_bits(t, n)
        for k, v in enumerate(FLAT_VALUES):
            want = o_svd.embed_plane_u8(planes[k], bits_row)
            assert np.array_equal(got[k], want), (v, row, np.unique(got[k].astype(int) - want))
            assert np.array_equal(marked_bits[k], o_svd.extract_plane(want)[0]), (v, row)
    for k, v in enumerate(FLAT_VALUES):
        assert np.array_equal(clean_bits[k], o_svd.extract_plane(planes[k])[0]), v


def test_piecewise_flat_planes_every_level_and_layout():
    """One flat tile per grey level 0..255 (twice over): uint8 aligned, uint8 unaligned (generic kernels) and the
    reference's float32 interleaved layout, embed and extract EQUAL the oracle (float32 within 1e-4, the
    synthesis rounding of idwt2)."""
    from b200wm import ops
    rng = np.random.RandomState(3)
    levels = np.concatenate([np.arange(256), rng.permutation(256)]).reshape(8, 64)
    plane = np.kron(levels, np.ones((8, 8))).astype(np.uint8)              # 64 x 512
    h, w = plane.shape
    n = h * w // 64
    wm = _wm((h, w))
    packed, nb = ops.pack_bits(wm[0], device=_dev())
    want = o_svd.embed_plane_u8(plane, wm[0])
    want_bits = o_svd.extract_plane(plane)[0]
    # aligned
    t = torch.from_numpy(plane.copy()).to(_dev())
    assert np.array_equal(_extract_bits(t, n)[0], want_bits)
    ops.dwtsvd_embed_(t, packed, nb)
    assert np.array_equal(t.cpu().numpy(), want)
    assert np.array_equal(_extract_bits(t, n)[0], o_svd.extract_plane(want)[0])
    # unaligned view
    big = torch.zeros((h + 2, w + 13), dtype=torch.uint8, device=_dev())
    view = big[1:h + 1, 5:w + 5]
    view.copy_(torch.from_numpy(plane))
    assert np.array_equal(_extract_bits(view, n)[0], want_bits)
    ops.dwtsvd_embed_(view, packed, nb)
    assert np.array_equal(view.cpu().numpy(), want)
    # float32 interleaved, channel 1 (negative and fractional flat values as well)
    yuv = np.zeros((h, w, 3), dtype=np.float32)
    yuv[:, :, 1] = plane.astype(np.float32) * 0.47 - 60.25
    want_f = o_svd.encode(yuv.copy(), wm)
    tf = torch.from_numpy(yuv.copy()).to(_dev())
    raw, _ = ops.dwtsvd_extract(tf, channel=1)
    assert np.array_equal(ops.unpack_bits(raw, n)[0], o_svd.decode(yuv)[0])
    ops.dwtsvd_embed_(tf, packed, nb, channel=1)
    np.testing.assert_allclose(tf.cpu().numpy()[:, :, 1], want_f[:, :, 1], rtol=0, atol=1e-4)
    raw, _ = ops.dwtsvd_extract(torch.from_numpy(want_f).to(_dev()), channel=1)
    assert np.array_equal(ops.unpack_bits(raw, n)[0], o_svd.decode(want_f)[0])
