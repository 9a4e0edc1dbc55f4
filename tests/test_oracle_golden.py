"""CPU: the oracle reproduces the reference's outputs stored in tests/golden (made by
oracle/make_golden.py from the reference itself)."""
import json
import os

import numpy as np
import pytest

from oracle import bracket, dct8, dwt_dct_svd as svd, haar, payload, synth
from parity import PAYLOAD, KEY


def _bits(packed, n):
    return np.unpackbits(packed)[:n].astype(np.float64).reshape(1, -1)


def test_haar_roundtrip_and_ll():
    rng = np.random.RandomState(0)
    x = rng.uniform(-100, 255, (64, 96)).astype(np.float32)
    ll, (lh, hl, hh) = haar.dwt2_haar(x)
    assert ll.dtype == np.float32 and ll.shape == (32, 48)
    s = (x[0::2, 0::2] + x[0::2, 1::2]) + (x[1::2, 0::2] + x[1::2, 1::2])
    np.testing.assert_allclose(ll, s / 2, rtol=3e-7, atol=1e-5)
    np.testing.assert_allclose(haar.idwt2_haar((ll, (lh, hl, hh))), x, rtol=0, atol=2e-4)


def test_dct_is_orthogonal_so_it_drops_out_of_the_svd_pair():
    """sigma(dct(B)) == sigma(B): the reason the CUDA path skips the 4x4 DCT (svd4.cuh)."""
    import cv2
    rng = np.random.RandomState(1)
    for _ in range(200):
        b = rng.uniform(-50, 500, (4, 4)).astype(np.float32)
        s_dct = np.linalg.svd(cv2.dct(b).astype(np.float64), compute_uv=False)
        s_raw = np.linalg.svd(b.astype(np.float64), compute_uv=False)
        np.testing.assert_allclose(s_dct, s_raw, rtol=1e-5, atol=1e-3)


def test_frame63_crop_dwtsvd(golden_dir):
    g = np.load(os.path.join(golden_dir, "frame63_crop.npz"))
    frame = g["bgr"]
    wm = payload.generate_wm(PAYLOAD, svd.wm_capacity(frame.shape), KEY)
    yuv = svd.encode(bracket.to_yuv(frame), wm)
    h, w = g["dwtsvd_marked_f32_ch1_window"].shape
    assert np.array_equal(yuv[:h, :w, 1], g["dwtsvd_marked_f32_ch1_window"])
    marked = bracket.from_yuv(yuv.copy())
    assert np.array_equal(marked.astype(np.int16) - frame, g["dwtsvd_marked_minus_src"])
    n = int(g["dwtsvd_nbits"])
    assert np.array_equal(svd.decode(bracket.to_yuv(frame)), _bits(g["dwtsvd_bits_clean"], n))
    assert np.array_equal(svd.decode(yuv), _bits(g["dwtsvd_bits_marked_f32"], n))
    assert np.array_equal(svd.decode(bracket.to_yuv(marked)), _bits(g["dwtsvd_bits_marked_u8"], n))
    for name in ("clean", "marked_f32", "marked_u8"):
        bits = _bits(g[f"dwtsvd_bits_{name}"], n)
        assert np.array_equal(payload.degenerate(bits, len(PAYLOAD), KEY), g[f"dwtsvd_pattern_{name}"])
    assert np.array_equal(g["dwtsvd_pattern_marked_u8"], PAYLOAD)


def test_frame63_crop_dct8(golden_dir):
    g = np.load(os.path.join(golden_dir, "frame63_crop.npz"))
    frame = g["bgr"]
    wm = payload.generate_wm(PAYLOAD, svd.wm_capacity(frame.shape), KEY)
    yuv = dct8.encode(bracket.to_yuv(frame), wm)
    h, w = g["dct8_marked_f32_ch1_window"].shape
    assert np.array_equal(yuv[:h, :w, 1], g["dct8_marked_f32_ch1_window"])
    marked = bracket.from_yuv(yuv.copy())
    assert np.array_equal(marked.astype(np.int16) - frame, g["dct8_marked_minus_src"])
    n = int(g["dct8_nbits"])
    assert np.array_equal(dct8.decode(yuv), _bits(g["dct8_bits_marked_f32"], n))
    assert np.array_equal(dct8.decode(bracket.to_yuv(marked)), _bits(g["dct8_bits_marked_u8"], n))


def test_in_mp4_frames(golden_dir):
    g = np.load(os.path.join(golden_dir, "in_mp4_frames.npz"))
    wm = g["wm"].astype(np.int64)
    for i in g["picks"]:
        frame = g[f"rgb_{i}"]
        marked = bracket.mark_frame(frame, lambda y: svd.encode(y, wm))
        assert np.array_equal(marked.astype(np.int16) - frame, g[f"marked_minus_src_{i}"])
        bits = svd.decode(bracket.to_yuv(marked))
        assert np.array_equal(bits, _bits(g[f"bits_{i}"], bits.size))
        assert np.array_equal(payload.degenerate(bits, 8, KEY), g["patterns_all_frames"][i])
    best, freq = payload.pattern_vote(list(g["patterns_all_frames"]))
    assert np.array_equal(best, g["vote_pattern"]) and freq == float(g["vote_frequency"])
    assert np.array_equal(best, PAYLOAD)


def test_synthetic_sizes(golden_dir):
    g = np.load(os.path.join(golden_dir, "synthetic_sizes.npz"))
    for n, (h, w) in enumerate(g["sizes"]):
        tag = f"{h}x{w}"
        if f"dwtsvd_{tag}_bits_marked" not in g:
            continue
        frame = synth.random_bgr(int(h), int(w), seed=100 + n)
        wm = payload.generate_wm(PAYLOAD, svd.wm_capacity(frame.shape), KEY)
        yuv = svd.encode(bracket.to_yuv(frame), wm)
        assert np.array_equal(yuv[:, :, 1], g[f"dwtsvd_{tag}_marked_ch1"])
        assert np.array_equal(svd.decode(yuv).reshape(-1), g[f"dwtsvd_{tag}_bits_marked"])
        assert np.array_equal(svd.decode(bracket.to_yuv(frame)).reshape(-1), g[f"dwtsvd_{tag}_bits_clean"])
        yuv8 = dct8.encode(bracket.to_yuv(frame), wm)
        assert np.array_equal(yuv8[:, :, 1], g[f"dct8_{tag}_marked_ch1"])
        assert np.array_equal(dct8.decode(yuv8).reshape(-1), g[f"dct8_{tag}_bits_marked"])


def test_per_block_form_equals_vectorised():
    frame = synth.random_bgr(40, 56, seed=5)
    wm = payload.generate_wm(PAYLOAD, svd.wm_capacity(frame.shape), KEY)
    a = svd.encode(bracket.to_yuv(frame), wm)
    b = svd.encode_per_block(bracket.to_yuv(frame), wm)
    assert np.array_equal(a, b)
    assert np.array_equal(svd.decode(a), svd.decode_per_block(b))


def test_payload_side(golden_dir):
    with open(os.path.join(golden_dir, "payload.json")) as f:
        pay = json.load(f)
    for k, perm in pay["permutations"].items():
        length, key = (int(v) for v in k.split(":"))
        assert payload.permutation(length, key).tolist() == perm
    assert payload.permutation(8, 0).tolist() == [6, 2, 1, 7, 3, 0, 5, 4]
    for case in pay["degenerate_cases"]:
        # rebuild a bit array with exactly these per-position counts
        length, n = case["length"], case["n"]
        bits = np.zeros(n)
        for i, c in enumerate(case["counts"]):
            idx = np.arange(i, n, length)
            bits[idx[:c]] = 1
        assert payload.degenerate(bits.reshape(1, -1), length, case["key"]).tolist() == case["pattern"]
    assert payload.payload_for_segment(5).tolist() == [0, 0, 0, 0, 0, 1, 0, 1]
    assert payload.payload_for_segment_copy(3, 2).tolist() == [0, 0, 1, 1, 0, 0, 1, 0]
    assert payload.segment_copy_from_pattern([0, 0, 1, 1, 0, 0, 1, 0]) == (3, 2)


def test_pattern_vote_ties_and_empty():
    a, b = np.array([0, 1]), np.array([1, 0])
    best, freq = payload.pattern_vote([b, a, a, b])
    assert best.tolist() == [1, 0] and freq == 0.5           # first seen wins the tie
    assert payload.pattern_vote([]) == (None, None)


def test_u8_plane_1080p(golden_dir):
    import hashlib
    g = np.load(os.path.join(golden_dir, "u8plane_1080p.npz"))
    y = synth.luma_plane_u8(1080, 1920, int(g["frame_index"]), int(g["seed"]))
    assert hashlib.sha256(y.tobytes()).hexdigest() == str(g["src_sha256"])
    bits = svd.extract_plane(y)
    assert np.array_equal(bits, _bits(g["bits_clean"], bits.size))


RESIZE_CASES = [((1080, 1920), (720, 1280)), ((240, 320), (160, 213)), ((64, 96), (48, 40)), ((37, 53), (20, 31)),
                ((48, 96), (24, 48)), ((48, 96), (16, 32)), ((48, 96), (24, 32)), ((30, 50), (30, 25)), ((9, 11), (1, 1))]


@pytest.mark.parametrize("src_hw,dst_hw", RESIZE_CASES)
def test_resize_restatement_is_cv2_bit_for_bit(src_hw, dst_hw):
    """The step-by-step restatement of cv2.resize that csrc/attacks.cu follows equals OpenCV itself."""
    import cv2
    from oracle import attacks
    rng = np.random.RandomState(src_hw[0] * 7 + dst_hw[1])
    src = rng.randint(0, 256, src_hw).astype(np.uint8)
    dsize = (dst_hw[1], dst_hw[0])
    small = cv2.resize(src, dsize, interpolation=cv2.INTER_AREA)
    assert np.array_equal(attacks.resize_area_restated(src, dsize), small)
    back = cv2.resize(small, (src_hw[1], src_hw[0]), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(attacks.resize_linear_restated(small, (src_hw[1], src_hw[0])), back)
    # bilinear reductions too (the kernel is not limited to enlarging)
    assert np.array_equal(attacks.resize_linear_restated(src, dsize), cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR))


def test_resize_restatement_random_geometries():
    """Seeded sweep over odd sizes and ratios (integer, 2x2, 1.5, irrational-looking): the restatement of
    cv2.resize that the CUDA kernels follow stays bit-identical to OpenCV."""
    import cv2
    from oracle import attacks
    rng = np.random.RandomState(2026)
    for _ in range(40):
        sh, sw = int(rng.randint(2, 90)), int(rng.randint(2, 120))
        dh, dw = int(rng.randint(1, sh + 1)), int(rng.randint(1, sw + 1))
        src = rng.randint(0, 256, (sh, sw)).astype(np.uint8)
        want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(attacks.resize_area_restated(src, (dw, dh)), want), ((sh, sw), (dh, dw))
        uh, uw = int(rng.randint(1, 140)), int(rng.randint(1, 160))
        want = cv2.resize(src, (uw, uh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(attacks.resize_linear_restated(src, (uw, uh)), want), ((sh, sw), (uh, uw))


def test_flat_blocks_are_exact_in_the_reference_restatement():
    """The rule the kernels use for flat tiles on a quantisation boundary (csrc/dwtsvd_tile.cuh:flat_sigma_ref):
    on a constant block the float32 Haar band is 4*fl(c*fl(c*v)) per coefficient and cv2.dct + LAPACK are exact,
    so sigma_0 = 16*fl(c*fl(c*|v|)) - for every uint8 level and for float samples (chroma can be negative)."""
    from oracle import dwt_dct_svd as o_svd
    c = np.float32(1.0 / np.sqrt(2.0))
    values = [np.float32(v) for v in range(256)] + list(np.random.RandomState(0).uniform(-100, 100, 200).astype(np.float32))
    for v in values:
        yuv = np.zeros((8, 8, 3), dtype=np.float32)
        yuv[:, :, 1] = v
        s32, _ = o_svd.decode_sigma(yuv)
        assert s32[0] == np.float32(16) * (c * (c * np.abs(v))), v
    # the case that motivated it: flat 255 reads as bit 1 (sigma_0 = 2039.99988, not 2040 = 136 * 15)
    assert o_svd.extract_plane(np.full((8, 8), 255, dtype=np.uint8))[0, 0] == 1.0
    assert o_svd.extract_plane(np.full((8, 8), 30, dtype=np.uint8))[0, 0] == 0.0


# ---- the reference's two media fixtures at full size (tests/golden/media, copied by oracle/make_golden.py) ----------
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_frame63(golden_dir):
    import cv2
    img = cv2.imread(os.path.join(golden_dir, "media", "frame63.jpeg"))
    g = np.load(os.path.join(golden_dir, "frame63_full.npz"))
    if img is None or _sha(img) != str(g["bgr_sha256"]):
        pytest.skip("this OpenCV build decodes frame63.jpeg differently from the one the goldens were made with")
    return img, g


def load_in_mp4(golden_dir):
    import cv2
    cap = cv2.VideoCapture(os.path.join(golden_dir, "media", "in.mp4"))
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(np.ascontiguousarray(f[:, :, ::-1]))     # BGR -> rgb24 byte order, like FileDecoder (frame_reader.py:59-63)
    cap.release()
    g = np.load(os.path.join(golden_dir, "in_mp4_all.npz"))
    if len(frames) != len(g["src_sha256"]) or any(_sha(f) != str(s) for f, s in zip(frames, g["src_sha256"])):
        pytest.skip("this OpenCV / FFmpeg build decodes in.mp4 differently from the one the goldens were made with")
    return frames, g


def test_oracle_equals_reference_on_whole_frame63(golden_dir):
    """frame63.jpeg at its full 1080p through both coder pairs: the oracle's marked frames hash to what the REFERENCE
    produced (oracle/make_golden.py ran it), its decoders return the reference's bits."""
    img, g = load_frame63(golden_dir)
    for tag, mod in (("dwtsvd", svd), ("dct8", dct8)):
        wm = payload.generate_wm(PAYLOAD, (1, img.shape[0] * img.shape[1] // 64), 0)
        yuv = mod.encode(bracket.to_yuv(img), wm)
        assert _sha(yuv[:, :, 1]) == str(g[f"{tag}_marked_f32_ch1_sha256"])
        marked = bracket.from_yuv(yuv.copy())
        assert _sha(marked) == str(g[f"{tag}_marked_u8_sha256"])
        n = int(g[f"{tag}_nbits"])
        for name, src in (("clean", bracket.to_yuv(img)), ("marked_f32", yuv), ("marked_u8", bracket.to_yuv(marked))):
            assert np.array_equal(mod.decode(src.copy())[0].astype(np.uint8), _bits(g[f"{tag}_bits_{name}"], n)[0]), (tag, name)
        assert np.array_equal(g[f"{tag}_pattern_marked_u8"], PAYLOAD)


def test_oracle_equals_reference_on_all_209_frames_of_in_mp4(golden_dir):
    """tests/mark.py + tests/detect.py on the reference's clip, every frame: marked frames hash to the reference's,
    raw bits and per-frame patterns are the reference's."""
    frames, g = load_in_mp4(golden_dir)
    wm = payload.generate_wm(PAYLOAD, (1, 240 * 320 // 64), 0)
    n = int(g["nbits"])
    for i, f in enumerate(frames):
        marked = bracket.mark_frame(f, lambda y: svd.encode(y, wm))
        assert _sha(marked) == str(g["marked_sha256"][i]), i
        bits = svd.decode(bracket.to_yuv(marked))
        assert np.array_equal(bits[0].astype(np.uint8), np.unpackbits(g["bits_marked"][i])[:n]), i
        assert np.array_equal(payload.degenerate(bits, 8, 0), g["patterns"][i]), i
