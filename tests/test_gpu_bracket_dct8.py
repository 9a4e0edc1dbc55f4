"""GPU: colour bracket and the 8x8 DCT pair vs the oracle (cv2 / reference restatement)."""
import os

import numpy as np
import pytest
import torch

from oracle import bracket, dct8 as o_dct, dwt_dct_svd as o_svd, payload as o_pay, synth
from parity import PAYLOAD, KEY

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_colour_bracket_is_bit_exact_vs_opencv():
    from b200wm import ops
    rng = np.random.RandomState(0)
    for shape in ((240, 320, 3), (128, 512, 3), (37, 53, 3), (1080, 1920, 3)):
        frame = rng.randint(0, 256, shape).astype(np.uint8)
        want = bracket.to_yuv(frame)
        got = ops.bgr8_to_yuv32(torch.from_numpy(frame).to(DEV)).cpu().numpy()
        w16 = shape[1] // 16 * 16          # OpenCV's scalar tail (width % vector) is not fused; compare the vector body
        assert np.array_equal(got[:, :w16], want[:, :w16])
        np.testing.assert_allclose(got, want, rtol=0, atol=3e-5)
        yuv = want + rng.normal(0, 3, shape).astype(np.float32)
        want8 = bracket.from_yuv(yuv.copy())
        got8 = ops.yuv32_to_bgr8(torch.from_numpy(yuv).to(DEV)).cpu().numpy()
        assert np.abs(got8[:, :w16].astype(np.int16) - want8[:, :w16]).max() == 0
        assert np.abs(got8.astype(np.int16) - want8).max() <= 1


def _masks_oracle(lum):
    coeffs = o_dct._dct_all(lum)
    return coeffs[..., 0, 0] / 8, o_dct.texture_mask_from_coeffs(coeffs), o_dct.luminance_mask_from_dc(coeffs[..., 0, 0])


def _frame_1080p(golden_dir):
    """The reference's 1080p still is not shipped whole (fixtures hold a 384x640 crop): tile the crop to 1080p and
    add seeded sensor-like noise so that the tiles are not copies of each other."""
    crop = np.load(os.path.join(golden_dir, "frame63_crop.npz"))["bgr"]
    reps = (-(-1080 // crop.shape[0]), -(-1920 // crop.shape[1]), 1)
    tile = np.tile(crop, reps)[:1080, :1920].astype(np.float32)
    return np.clip(np.rint(tile + np.random.RandomState(63).normal(0, 1.5, tile.shape)), 0, 255).astype(np.uint8)


def _sources(golden_dir):
    yield "frame63 crop", np.load(os.path.join(golden_dir, "frame63_crop.npz"))["bgr"]
    yield "synthetic 128x192", synth.random_bgr(128, 192, 21)
    yield "synthetic 100x132 (ragged)", synth.random_bgr(100, 132, 9)
    yield "1080p (fixture tiled + noise)", _frame_1080p(golden_dir)


def _explained(diff_blocks, ties, what, limit=5e-3):
    """Every block in ``diff_blocks`` must carry one of the tie reasons; and ties may excuse only a small share."""
    unexplained = diff_blocks & ~ties
    assert not unexplained.any(), f"{what}: {int(unexplained.sum())} blocks differ without a float32 tie, e.g. {np.argwhere(unexplained)[:4].tolist()}"
    assert diff_blocks.mean() < limit, f"{what}: {diff_blocks.mean():.2e} of the blocks differ (all ties, but too many)"
    return int(diff_blocks.sum())


def test_dct8_masks_vs_oracle(golden_dir):
    """dct_encoder.py:41-102: block means within float32 rounding, frame sum, and the texture mask EQUAL to the
    reference's wherever no comparison of its decision tree is a float32 tie (oracle/dct8.py:tie_blocks names them)."""
    from b200wm import ops
    for name, frame in _sources(golden_dir):
        yuv = bracket.to_yuv(frame)
        mean_o, tex_o, lum_o = _masks_oracle(yuv[:, :, 0])
        block_mean, tex, frame_sum = ops.dct8_masks(torch.from_numpy(yuv).to(DEV), channel=0)
        np.testing.assert_allclose(block_mean[0].cpu().numpy(), mean_o.reshape(-1), rtol=2e-6, atol=2e-5)
        np.testing.assert_allclose(float(frame_sum[0]), mean_o.astype(np.float64).sum(), rtol=1e-6)
        tex_g = tex[0].cpu().numpy().reshape(tex_o.shape)
        differ = ~np.isclose(tex_g, tex_o, rtol=2e-5, atol=1e-6)        # the ramp 1 + 1.25 (eh - 290) / 1510 is continuous in eh
        n = _explained(differ, o_dct.tie_blocks(yuv)["mask"], f"texture mask, {name}")
        print(f"{name}: texture mask differs on {n} of {differ.size} blocks, all decision-tree ties")


def test_public_mask_methods_of_the_dct_plugins(golden_dir):
    """``DctEncoder.luminance_mask`` / ``texture_mask`` (dct_encoder.py:41-102; the decoder carries copies,
    dct_decoder.py:29-89) are public in the reference: same arrays here, float64 ``[H/8, W/8]``, numpy in -> numpy out,
    strided views of an interleaved frame accepted as the reference passes them (``yuv[:, :, 0]``)."""
    from offmark_b200.embed.dct_encoder import DctEncoder
    from offmark_b200.extract.dct_decoder import DctDecoder
    for name, frame in list(_sources(golden_dir))[:3]:
        yuv = bracket.to_yuv(frame)
        _, tex_o, lum_o = _masks_oracle(yuv[:, :, 0])
        for plugin in (DctEncoder(), DctDecoder()):
            lum = plugin.luminance_mask(yuv[:, :, 0])
            tex = plugin.texture_mask(yuv[:, :, 0])
            assert isinstance(lum, np.ndarray) and lum.dtype == np.float64 and lum.shape == lum_o.shape == tex.shape
            # the brightness rule is continuous but for its two steps at 15 and 25: a block mean within float32 rounding
            # of a step may land on the other side
            near_step = np.isclose(_masks_oracle(yuv[:, :, 0])[0], 15, atol=1e-4) | np.isclose(_masks_oracle(yuv[:, :, 0])[0], 25, atol=1e-4)
            assert np.allclose(lum[~near_step], lum_o[~near_step], rtol=1e-6, atol=1e-6), name
            _explained(~np.isclose(tex, tex_o, rtol=2e-5, atol=1e-6), o_dct.tie_blocks(yuv)["mask"], f"texture_mask(), {name}")
        on_device = DctEncoder().texture_mask(torch.from_numpy(yuv).to(DEV)[:, :, 0])
        assert on_device.is_cuda and np.array_equal(on_device.cpu().numpy(), tex)
    with pytest.raises(ValueError):
        DctEncoder().luminance_mask(np.zeros((16, 16)))


@pytest.mark.parametrize("source", ["frame63 crop", "synthetic 128x192", "synthetic 100x132 (ragged)", "1080p (fixture tiled + noise)"])
def test_dct8_embed_extract_vs_oracle(golden_dir, source):
    """DctEncoder.encode / DctDecoder.decode on the reference's float32 interleaved layout: marked channel within 2e-3
    (float32) of the reference on EVERY block that is not a nameable float32 tie (mask decision, sign of a c21 below
    noise, floor or rounding on a half-step), raw bits likewise, golden bit arrays of the reference itself, votes."""
    from b200wm import ops
    frame = dict(_sources(golden_dir))[source]
    yuv0 = bracket.to_yuv(frame)
    wm = o_pay.generate_wm(PAYLOAD, o_svd.wm_capacity(frame.shape), KEY)
    want = o_dct.encode(yuv0.copy(), wm)
    t = torch.from_numpy(yuv0).to(DEV)
    packed, n = ops.pack_bits(wm[0], device=DEV)
    masks = ops.dct8_masks(t, channel=0)
    ops.dct8_embed_(t, masks, packed, n, alpha=20, channel=1)
    got = t.cpu().numpy()
    assert np.array_equal(got[:, :, 0], yuv0[:, :, 0]) and np.array_equal(got[:, :, 2], yuv0[:, :, 2])
    by, bx = frame.shape[0] // 8, frame.shape[1] // 8
    assert np.array_equal(got[by * 8:, :, 1], yuv0[by * 8:, :, 1]) and np.array_equal(got[:, bx * 8:, 1], yuv0[:, bx * 8:, 1])
    ties = o_dct.tie_blocks(yuv0)
    err = np.abs(got[:by * 8, :bx * 8, 1] - want[:by * 8, :bx * 8, 1]).reshape(by, 8, bx, 8).max(axis=(1, 3))
    n_embed = _explained(err >= 2e-3, ties["mask"] | ties["sign"] | ties["floor"], f"embed, {source}", limit=6e-2)
    # after the uint8 bracket (video/embedder.py:36-38) the frames agree within 1 LSB on those same blocks
    d8 = np.abs(bracket.from_yuv(got.copy()).astype(np.int16) - bracket.from_yuv(want.copy())).max(axis=2)
    d8 = d8[:by * 8, :bx * 8].reshape(by, 8, bx, 8).max(axis=(1, 3))
    _explained(d8 > 1, ties["mask"] | ties["sign"] | ties["floor"], f"embed after the uint8 bracket, {source}", limit=6e-2)
    # extraction: our extractor on the reference's marked frame (and on the clean one), against the reference's
    for label, src in (("marked", want), ("clean", yuv0)):
        bits_ref = o_dct.decode(src.copy())
        ts = torch.from_numpy(src).to(DEV)
        raw, counts = ops.dct8_extract(ts, ops.dct8_masks(ts, channel=0), alpha=20, payload_len=8, channel=1)
        bits = ops.unpack_bits(raw, bits_ref.size)
        t2 = o_dct.tie_blocks(src)
        differ = np.zeros(by * bx, dtype=bool) | (bits[0, :by * bx] != bits_ref[0, :by * bx])
        n_x = _explained(differ.reshape(by, bx), t2["mask"] | t2["round"], f"extract {label}, {source}", limit=1e-2)
        assert not bits[0, by * bx:].any()
        assert np.array_equal(o_pay.degenerate(bits.astype(np.float64), 8, KEY), o_pay.degenerate(bits_ref, 8, KEY))
        assert counts[0].cpu().tolist() == [int(bits[0][i::8].sum()) for i in range(8)]
        print(f"{source}: extract {label}: {n_x} of {by * bx} bits differ, all ties; embed: {n_embed} blocks differ, all ties")
    assert np.array_equal(o_pay.degenerate(o_dct.decode(got.copy()), 8, KEY), PAYLOAD)
    if source == "frame63 crop":        # the reference's own outputs (oracle/make_golden.py ran the reference itself)
        g = np.load(os.path.join(golden_dir, "frame63_crop.npz"))
        nb = int(g["dct8_nbits"])
        ref_u8 = (frame.astype(np.int16) + g["dct8_marked_minus_src"]).astype(np.uint8)
        for key, src in (("dct8_bits_clean", yuv0), ("dct8_bits_marked_f32", want), ("dct8_bits_marked_u8", bracket.to_yuv(ref_u8))):
            ts = torch.from_numpy(src).to(DEV)
            raw, _ = ops.dct8_extract(ts, ops.dct8_masks(ts, channel=0), alpha=20, channel=1)
            gold = np.unpackbits(g[key])[:nb]
            t2 = o_dct.tie_blocks(src)
            differ = (ops.unpack_bits(raw, nb)[0, :by * bx] != gold[:by * bx]).reshape(by, bx)
            _explained(differ, t2["mask"] | t2["round"], f"golden {key}", limit=1e-2)
        h, w = g["dct8_marked_f32_ch1_window"].shape
        werr = np.abs(got[:h, :w, 1] - g["dct8_marked_f32_ch1_window"]).reshape(h // 8, 8, w // 8, 8).max(axis=(1, 3))
        _explained(werr >= 2e-3, (ties["mask"] | ties["sign"] | ties["floor"])[:h // 8, :w // 8], "golden marked window", limit=6e-2)


def test_dct8_planar_u8_1080p_vs_oracle(golden_dir):
    """The same pair on planar uint8 4:4:4 planes at 1080p (masks from Y, mark in U): marked plane within 1 LSB of
    the reference flow (float YUV -> DctEncoder -> clip / around, video/embedder.py:37-38) on every block that is not
    a nameable tie, raw bits of the marked plane likewise, identical votes."""
    from b200wm import ops
    frame = _frame_1080p(golden_dir)
    yp, up = np.ascontiguousarray(frame[:, :, 1]), np.ascontiguousarray(frame[:, :, 0])
    h, w = yp.shape
    wm = o_pay.generate_wm(PAYLOAD, (1, h * w // 64), KEY)
    packed, n = ops.pack_bits(wm[0], device=DEV)
    yuv = np.zeros((h, w, 3), dtype=np.float32)
    yuv[:, :, 0], yuv[:, :, 1] = yp, up
    ties = o_dct.tie_blocks(yuv)
    want = np.around(np.clip(o_dct.encode(yuv.copy(), wm)[:, :, 1], 0, 255)).astype(np.uint8)
    ty, tu = torch.from_numpy(yp).to(DEV), torch.from_numpy(up.copy()).to(DEV)
    masks = ops.dct8_masks(ty)
    ops.dct8_embed_(tu, masks, packed, n, alpha=20)
    got = tu.cpu().numpy()
    by, bx = h // 8, w // 8
    d = np.abs(got.astype(np.int16) - want).reshape(by, 8, bx, 8).max(axis=(1, 3))
    n_embed = _explained(d > 1, ties["mask"] | ties["sign"] | ties["floor"], "planar u8 embed 1080p", limit=6e-2)   # tolerance: 1 LSB
    yuv[:, :, 1] = got
    bits_ref = o_dct.decode(yuv)
    raw, counts = ops.dct8_extract(tu, masks, alpha=20, payload_len=8)
    bits = ops.unpack_bits(raw, h * w // 64)
    t2 = o_dct.tie_blocks(yuv)
    n_x = _explained((bits[0] != bits_ref[0]).reshape(by, bx), t2["mask"] | t2["round"], "planar u8 extract 1080p", limit=1e-2)
    assert np.array_equal(o_pay.degenerate(bits.astype(np.float64), 8, KEY), o_pay.degenerate(bits_ref, 8, KEY))
    assert np.array_equal(o_pay.degenerate(bits.astype(np.float64), 8, KEY), PAYLOAD)
    print(f"1080p planar u8: {n_embed} blocks off by more than 1 LSB and {n_x} bits differ of {by * bx}, all ties")


def test_dct8_uint8_planes_fast_path_vs_generic_and_oracle():
    """Planar uint8 (yuv444 planes): the vector-load instantiations (8-byte aligned rows) give the same
    masks, bytes and bits as the generic byte-wise ones (unaligned view), and the marked plane equals the
    reference flow on the same values (float YUV -> DctEncoder -> clip/around, dct_encoder.py:18-39 +
    video/embedder.py:37-38) within 1 LSB wherever the sign of c21 is not float32 noise."""
    from b200wm import ops
    h, w = 136, 200
    rng = np.random.RandomState(77)
    base = synth.random_bgr(h, w, 5)
    yp = base[:, :, 0].copy()
    up = np.clip(base[:, :, 1].astype(np.int16) + rng.randint(-3, 4, (h, w)), 0, 255).astype(np.uint8)
    up[:16, :16] = 255                      # saturated and flat blocks (c21 == 0 exactly: stay unmarked)
    up[16:24, :8] = 0
    wm = o_pay.generate_wm(PAYLOAD, (1, h * w // 64), KEY)
    packed, n = ops.pack_bits(wm[0], device=DEV)

    def run(aligned):
        if aligned:
            ty, tu = torch.from_numpy(yp.copy()).to(DEV), torch.from_numpy(up.copy()).to(DEV)
        else:
            big = torch.zeros((2, h + 1, w + 11), dtype=torch.uint8, device=DEV)
            ty, tu = big[0, 1:, 3:w + 3], big[1, 1:, 3:w + 3]
            ty.copy_(torch.from_numpy(yp)); tu.copy_(torch.from_numpy(up))
        masks = ops.dct8_masks(ty)
        ops.dct8_embed_(tu, masks, packed, n, alpha=20)
        raw, counts = ops.dct8_extract(tu, masks, alpha=20, payload_len=8)
        return [m.cpu().numpy() for m in masks], tu.cpu().numpy(), raw.cpu().numpy(), counts.cpu().numpy()
    m_a, u_a, raw_a, cnt_a = run(True)
    m_g, u_g, raw_g, cnt_g = run(False)
    for a, b in zip(m_a, m_g):
        assert np.array_equal(a, b)
    assert np.array_equal(u_a, u_g) and np.array_equal(raw_a, raw_g) and np.array_equal(cnt_a, cnt_g)
    # reference flow on the same values
    yuv = np.zeros((h, w, 3), dtype=np.float32)
    yuv[:, :, 0], yuv[:, :, 1] = yp, up
    want = np.around(np.clip(o_dct.encode(yuv.copy(), wm)[:, :, 1], 0, 255)).astype(np.uint8)
    by, bx = h // 8, w // 8
    c21 = o_dct._dct_all(yuv[:, :, 1])[..., 2, 1]
    err = np.abs(u_a[:by * 8, :bx * 8].astype(np.int16) - want[:by * 8, :bx * 8]).reshape(by, 8, bx, 8).max(axis=(1, 3))
    solid = np.abs(c21) > 1e-3
    assert (err[solid] <= 1).mean() > 0.995, (err[solid] <= 1).mean()        # tolerance: 1 LSB
    assert np.array_equal(u_a[by * 8:], up[by * 8:]) and np.array_equal(u_a[:, bx * 8:], up[:, bx * 8:])
    # and the payload reads back, by us and by the reference decoder
    bits = ops.unpack_bits(torch.from_numpy(raw_a), h * w // 64)
    assert np.array_equal(o_pay.degenerate(bits.astype(np.float64), 8, KEY), PAYLOAD)
    yuv[:, :, 1] = u_a
    assert np.array_equal(o_pay.degenerate(o_dct.decode(yuv), 8, KEY), PAYLOAD)


def test_dct8_fused_calls_equal_the_three_kernel_path(golden_dir):
    """b200wm_dct8_encode / _decode (the pair as the reference calls it: masks + quantiser per call, no caller-visible
    mask arrays) give the very planes, raw bits and counts of b200wm_dct8_masks + _embed / _extract - on the reference's
    float32 interleaved layout, planar uint8 (vector path) and unaligned planar uint8 (generic path), batched with
    per-frame watermark rows - and the plugin classes, which use them, still match the oracle."""
    from b200wm import ops
    rng = np.random.RandomState(5)
    frames = np.stack([synth.random_bgr(136, 200, s) for s in range(3)])
    frames[1, :24, :32] = 200                                   # flat blocks: c21 == 0 exactly, stay unmarked
    n_bits = 136 * 200 // 64
    rows = rng.randint(0, 2, (2, n_bits))
    packed, n = ops.pack_bits(rows, device=DEV)
    frame_row = torch.tensor([0, 1, 0], dtype=torch.int32, device=DEV)
    # float32 interleaved
    yuv = np.stack([bracket.to_yuv(f) for f in frames])
    a, b = torch.from_numpy(yuv.copy()).to(DEV), torch.from_numpy(yuv.copy()).to(DEV)
    masks = ops.dct8_masks(a, channel=0)
    ops.dct8_embed_(a, masks, packed, n, alpha=20, channel=1, frame_wm_row=frame_row)
    ops.dct8_encode_(b, b, packed, n, alpha=20, lum_channel=0, channel=1, frame_wm_row=frame_row)
    assert torch.equal(a, b)
    raw_a, cnt_a = ops.dct8_extract(a, ops.dct8_masks(a, channel=0), alpha=20, payload_len=8, channel=1)
    raw_b, cnt_b = ops.dct8_decode(b, b, alpha=20, payload_len=8, lum_channel=0, channel=1)
    assert torch.equal(raw_a, raw_b) and torch.equal(cnt_a, cnt_b)
    # planar uint8, aligned and unaligned
    for aligned in (True, False):
        def planes(c):
            src = np.ascontiguousarray(frames[:, :, :, c])
            if aligned:
                return torch.from_numpy(src.copy()).to(DEV)
            big = torch.zeros((3, 137, 211), dtype=torch.uint8, device=DEV)
            view = big[:, 1:, 3:203]
            view.copy_(torch.from_numpy(src))
            return view
        ya, ua, ub = planes(1), planes(0), planes(0)
        masks = ops.dct8_masks(ya)
        ops.dct8_embed_(ua, masks, packed, n, alpha=20, frame_wm_row=frame_row)
        ops.dct8_encode_(ya, ub, packed, n, alpha=20, frame_wm_row=frame_row)
        assert torch.equal(ua, ub), aligned
        raw_a, cnt_a = ops.dct8_extract(ua, masks, alpha=20, payload_len=8)
        raw_b, cnt_b = ops.dct8_decode(ya, ub, alpha=20, payload_len=8)
        assert torch.equal(raw_a, raw_b) and torch.equal(cnt_a, cnt_b), aligned
    with pytest.raises(ValueError):                              # the two channels must share dtype and geometry
        ops.dct8_decode(torch.zeros((1, 64, 64), dtype=torch.uint8, device=DEV), torch.zeros((1, 64, 72), dtype=torch.uint8, device=DEV))
    with pytest.raises(IndexError):
        ops.dct8_encode_(ya, ub, packed[:, :1].contiguous(), 32)
