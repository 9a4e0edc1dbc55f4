"""Generate tests/golden/* by running the REFERENCE ITSELF, and pin the oracle to it.

Test infrastructure (see oracle/__init__.py).  Runs only in the build
container, where /root/reference exists:

    python -m oracle.make_golden            # from the repo root

It imports the reference's modules verbatim from /root/reference/src (the one
missing dependency, PyWavelets, is replaced by the Haar stand-in in
oracle/_shim, see oracle/haar.py), feeds them the reference's own fixtures
(tests/media/imgs/frame63.jpeg, tests/media/in.mp4) and seeded synthetic
frames, asserts that every oracle function reproduces the reference's output
BIT FOR BIT on those inputs, and writes the inputs/outputs that the tests need
as small fixtures.  Nothing here travels to the GPU box except the fixtures.
"""
import hashlib
import json
import os
import sys

import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, os.path.join(REF, "src"))

from offmark.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder      # noqa: E402  (reference)
from offmark.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder    # noqa: E402
from offmark.embed.dct_encoder import DctEncoder                    # noqa: E402
from offmark.extract.dct_decoder import DctDecoder                  # noqa: E402
from offmark.generator.shuffler import Shuffler                     # noqa: E402
from offmark.generator.grayscale import GrayScale                   # noqa: E402
from offmark.degenerator.de_shuffler import DeShuffler              # noqa: E402
from offmark.degenerator.de_grayscale import DeGrayScale            # noqa: E402

from oracle import dwt_dct_svd as o_svd                             # noqa: E402
from oracle import dct8 as o_dct                                    # noqa: E402
from oracle import payload as o_pay                                 # noqa: E402
from oracle import bracket as o_br                                  # noqa: E402
from oracle import synth                                            # noqa: E402

PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])      # tests/mark.py:22
KEY = 0


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b), f"oracle != reference: {what} ({np.count_nonzero(a != b)} differ)"
    print(f"  ok  oracle == reference  {what}")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_mark_frame(frame_u8, encoder):
    """The body of src/offmark/video/embedder.py:33-39 driven with the reference encoder."""
    yuv = cv2.cvtColor(frame_u8.astype(np.float32), cv2.COLOR_BGR2YUV)
    yuv = encoder.encode(yuv)
    out = cv2.cvtColor(yuv, cv2.COLOR_YUV2BGR)
    return np.around(np.clip(out, a_min=0, a_max=255)).astype(np.uint8)


def ref_check_frame(frame_u8, decoder, degenerator):
    yuv = cv2.cvtColor(frame_u8.astype(np.float32), cv2.COLOR_BGR2YUV)
    bits = decoder.decode(yuv)
    return bits, degenerator.degenerate(bits)


def run_pair(frame_u8, tag, enc_cls, dec_cls, o_encode, o_decode, fixtures, f32_window=None):
    """Reference vs oracle on one uint8 frame for one encoder/decoder pair."""
    enc, dec = enc_cls(), dec_cls()
    wm = Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(frame_u8.shape))
    same(o_pay.generate_wm(PAYLOAD, enc.wm_capacity(frame_u8.shape), KEY), wm, f"{tag} generate_wm")
    enc.read_wm(wm)
    deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)

    yuv0 = cv2.cvtColor(frame_u8.astype(np.float32), cv2.COLOR_BGR2YUV)
    same(o_br.to_yuv(frame_u8), yuv0, f"{tag} bracket to_yuv")

    yuv_ref = enc.encode(yuv0.copy())
    yuv_orc = o_encode(yuv0.copy(), wm)
    same(yuv_orc, yuv_ref, f"{tag} encode (float32 yuv)")

    marked_ref = ref_mark_frame(frame_u8, enc)
    same(o_br.from_yuv(yuv_ref.copy()), marked_ref, f"{tag} bracket from_yuv")

    out = {}
    for name, src in (("clean", yuv0), ("marked_f32", yuv_ref),
                      ("marked_u8", cv2.cvtColor(marked_ref.astype(np.float32), cv2.COLOR_BGR2YUV))):
        bits_ref = dec.decode(src.copy())
        same(o_decode(src.copy()), bits_ref, f"{tag} decode {name}")
        pat_ref = deg.degenerate(bits_ref)
        same(o_pay.degenerate(bits_ref, len(PAYLOAD), KEY), pat_ref, f"{tag} degenerate {name}")
        out[f"bits_{name}"] = np.packbits(bits_ref.astype(np.uint8).reshape(-1))
        out[f"pattern_{name}"] = pat_ref
    assert np.array_equal(out["pattern_marked_f32"], PAYLOAD)
    out["nbits"] = np.int64(bits_ref.size)
    out["marked_minus_src"] = (marked_ref.astype(np.int16) - frame_u8.astype(np.int16)).astype(np.int8)
    assert np.array_equal(frame_u8.astype(np.int16) + out["marked_minus_src"], marked_ref)
    if f32_window is not None:
        h, w = f32_window
        out["marked_f32_ch1_window"] = yuv_ref[:h, :w, 1].copy()
    out["marked_f32_ch1_sha256"] = np.array(sha(yuv_ref[:, :, 1]))
    for k, v in out.items():
        fixtures[f"{tag}_{k}"] = v


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {"numpy": np.__version__, "cv2": cv2.__version__,
            "pywt": "Haar stand-in restating PyWavelets 1.4.1 (oracle/haar.py)",
            "payload": PAYLOAD.tolist(), "key": KEY}

    # ---------------------------------------------------------------- frame63 crop
    print("frame63.jpeg crop (reference fixture tests/media/imgs/frame63.jpeg)")
    img = cv2.imread(os.path.join(REF, "tests", "media", "imgs", "frame63.jpeg"))
    assert img is not None and img.shape == (1080, 1920, 3)
    crop = np.ascontiguousarray(img[300:300 + 384, 640:640 + 640])
    fx = {"bgr": crop}
    run_pair(crop, "dwtsvd", DwtDctSvdEncoder, DwtDctSvdDecoder, o_svd.encode, o_svd.decode, fx, (128, 256))
    run_pair(crop, "dct8", DctEncoder, DctDecoder, o_dct.encode, o_dct.decode, fx, (128, 256))
    np.savez_compressed(os.path.join(OUT, "frame63_crop.npz"), **fx)

    # ---------------------------------------------------------------- in.mp4 frames
    print("in.mp4 frames (reference fixture tests/media/in.mp4, read as rgb24 like FileDecoder)")
    cap = cv2.VideoCapture(os.path.join(REF, "tests", "media", "in.mp4"))
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(np.ascontiguousarray(f[:, :, ::-1]))     # BGR -> rgb24 byte order
    cap.release()
    assert len(frames) == 209 and frames[0].shape == (240, 320, 3)
    meta["in_mp4_frames"] = len(frames)
    picks = [0, 70, 140, 208]
    enc, dec = DwtDctSvdEncoder(), DwtDctSvdDecoder()
    wm = Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(frames[0].shape))
    enc.read_wm(wm)
    deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
    fx = {"picks": np.array(picks), "wm": wm.astype(np.uint8)}
    # whole clip through the reference flow (mark.py + detect.py in memory): per-frame patterns
    patterns = []
    for i, f in enumerate(frames):
        marked = ref_mark_frame(f, enc)
        bits, pat = ref_check_frame(marked, dec, deg)
        patterns.append(pat)
        if i in picks:
            same(o_br.mark_frame(f, lambda y: o_svd.encode(y, wm)), marked, f"in.mp4[{i}] mark_frame")
            same(o_svd.decode(o_br.to_yuv(marked)), bits, f"in.mp4[{i}] decode")
            fx[f"rgb_{i}"] = f
            fx[f"marked_minus_src_{i}"] = (marked.astype(np.int16) - f.astype(np.int16)).astype(np.int8)
            fx[f"bits_{i}"] = np.packbits(bits.astype(np.uint8).reshape(-1))
    patterns = np.array(patterns)
    fx["patterns_all_frames"] = patterns
    best, freq = o_pay.pattern_vote(list(patterns))
    print("  clip vote:", best, freq)
    assert np.array_equal(best, PAYLOAD)
    fx["vote_pattern"], fx["vote_frequency"] = best, np.float64(freq)
    np.savez_compressed(os.path.join(OUT, "in_mp4_frames.npz"), **fx)

    # ---------------------------------------------------------------- the reference's two media fixtures, whole
    # The two files are small (0.4 MB + 0.3 MB) test DATA of the reference, copied next to the goldens so that the GPU
    # box - which has no /root/reference - can run the very inputs the reference's own scripts use at full size:
    # frame63.jpeg at 1080p and all 209 frames of in.mp4.  What the reference made of them is stored as hashes, packed
    # bits and patterns; the oracle reproduces the marked frames on the box and the hashes prove it is still the
    # reference's output there.
    import shutil
    media = os.path.join(OUT, "media")
    os.makedirs(media, exist_ok=True)
    meta["media"] = {}
    for rel in (("imgs", "frame63.jpeg"), ("in.mp4",)):
        src_path = os.path.join(REF, "tests", "media", *rel)
        shutil.copyfile(src_path, os.path.join(media, rel[-1]))
        with open(src_path, "rb") as fh:
            meta["media"][rel[-1]] = {"from": "tests/media/" + "/".join(rel), "sha256": hashlib.sha256(fh.read()).hexdigest()}
    print("frame63.jpeg, whole 1080p frame, both coder pairs")
    fx = {"bgr_sha256": np.array(sha(img))}
    for tag, enc_cls, dec_cls, o_e, o_d in (("dwtsvd", DwtDctSvdEncoder, DwtDctSvdDecoder, o_svd.encode, o_svd.decode),
                                            ("dct8", DctEncoder, DctDecoder, o_dct.encode, o_dct.decode)):
        enc, dec = enc_cls(), dec_cls()
        wm = Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(img.shape))
        enc.read_wm(wm)
        deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
        yuv0 = cv2.cvtColor(img.astype(np.float32), cv2.COLOR_BGR2YUV)
        yuv_ref = enc.encode(yuv0.copy())
        same(o_e(yuv0.copy(), wm), yuv_ref, f"frame63 full {tag} encode")
        marked = ref_mark_frame(img, enc)
        fx[f"{tag}_marked_f32_ch1_sha256"] = np.array(sha(yuv_ref[:, :, 1]))
        fx[f"{tag}_marked_u8_sha256"] = np.array(sha(marked))
        for name, src in (("clean", yuv0), ("marked_f32", yuv_ref),
                          ("marked_u8", cv2.cvtColor(marked.astype(np.float32), cv2.COLOR_BGR2YUV))):
            bits = dec.decode(src.copy())
            same(o_d(src.copy()), bits, f"frame63 full {tag} decode {name}")
            fx[f"{tag}_bits_{name}"] = np.packbits(bits.astype(np.uint8).reshape(-1))
            fx[f"{tag}_pattern_{name}"] = deg.degenerate(bits)
        fx[f"{tag}_nbits"] = np.int64(bits.size)
    np.savez_compressed(os.path.join(OUT, "frame63_full.npz"), **fx)

    print("in.mp4, all 209 frames through mark.py + detect.py (in memory)")
    enc, dec = DwtDctSvdEncoder(), DwtDctSvdDecoder()
    wm = Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity(frames[0].shape))
    enc.read_wm(wm)
    deg = DeShuffler(key=KEY).set_shape(PAYLOAD.shape)
    src_sha, marked_sha, bits_all, pats = [], [], [], []
    for i, f in enumerate(frames):
        marked = ref_mark_frame(f, enc)
        bits, pat = ref_check_frame(marked, dec, deg)
        src_sha.append(sha(f))
        marked_sha.append(sha(marked))
        bits_all.append(np.packbits(bits.astype(np.uint8).reshape(-1)))
        pats.append(pat)
        if i % 19 == 0:
            same(o_br.mark_frame(f, lambda y: o_svd.encode(y, wm)), marked, f"in.mp4[{i}] mark_frame (all-frames pass)")
    assert np.array_equal(np.array(pats), patterns)
    np.savez_compressed(os.path.join(OUT, "in_mp4_all.npz"), src_sha256=np.array(src_sha), marked_sha256=np.array(marked_sha),
                        bits_marked=np.stack(bits_all), patterns=np.array(pats), nbits=np.int64(bits.size))

    # ---------------------------------------------------------------- synthetic float yuv, odd sizes
    print("seeded synthetic float32 frames, sizes that exercise the truncation rules")
    fx = {}
    sizes = [(8, 8), (7, 9), (16, 24), (37, 53), (70, 90), (100, 132), (64, 64)]
    for n, (h, w) in enumerate(sizes):
        frame = synth.random_bgr(h, w, seed=100 + n)
        tag = f"{h}x{w}"
        for pair, enc_cls, dec_cls, o_e, o_d in (("dwtsvd", DwtDctSvdEncoder, DwtDctSvdDecoder, o_svd.encode, o_svd.decode),
                                                 ("dct8", DctEncoder, DctDecoder, o_dct.encode, o_dct.decode)):
            enc, dec = enc_cls(), dec_cls()
            cap_shape = enc.wm_capacity(frame.shape)
            if cap_shape[1] == 0:
                continue
            wm = Shuffler(key=KEY).generate_wm(PAYLOAD, cap_shape)
            enc.read_wm(wm)
            yuv0 = o_br.to_yuv(frame)
            if pair == "dct8" and (h % 8 or w % 8):
                # reference DctEncoder needs whole 8x8 blocks for its capacity to cover the walk
                pass
            yuv_ref = enc.encode(yuv0.copy())
            same(o_e(yuv0.copy(), wm), yuv_ref, f"{pair} {tag} encode")
            bits = dec.decode(yuv_ref.copy())
            same(o_d(yuv_ref.copy()), bits, f"{pair} {tag} decode marked")
            bits0 = dec.decode(yuv0.copy())
            same(o_d(yuv0.copy()), bits0, f"{pair} {tag} decode clean")
            fx[f"{pair}_{tag}_bits_marked"] = bits.astype(np.uint8).reshape(-1)
            fx[f"{pair}_{tag}_bits_clean"] = bits0.astype(np.uint8).reshape(-1)
            fx[f"{pair}_{tag}_marked_ch1"] = yuv_ref[:, :, 1].copy()
    fx["sizes"] = np.array(sizes)
    np.savez_compressed(os.path.join(OUT, "synthetic_sizes.npz"), **fx)

    # ---------------------------------------------------------------- planar uint8 luma, 1080p (north_star layout)
    print("planar uint8 1080p luma plane through the reference (plane packed into channel 1)")
    y = synth.luma_plane_u8(1080, 1920, frame_index=0, seed=20261018)
    enc, dec = DwtDctSvdEncoder(), DwtDctSvdDecoder()
    wm = Shuffler(key=KEY).generate_wm(PAYLOAD, enc.wm_capacity((1080, 1920, 3)))
    enc.read_wm(wm)
    yuv = np.zeros((1080, 1920, 3), dtype=np.float32)
    yuv[:, :, 1] = y
    bits_clean = dec.decode(yuv.copy())
    same(o_svd.extract_plane(y), bits_clean, "u8 plane extract clean")
    enc.encode(yuv)
    marked = np.around(np.clip(yuv[:, :, 1], 0, 255)).astype(np.uint8)
    same(o_svd.embed_plane_u8(y, wm[0]), marked, "u8 plane embed")
    yuv[:, :, 1] = marked
    bits_marked = dec.decode(yuv.copy())
    same(o_svd.extract_plane(marked), bits_marked, "u8 plane extract marked")
    np.savez_compressed(
        os.path.join(OUT, "u8plane_1080p.npz"),
        seed=np.int64(20261018), frame_index=np.int64(0),
        src_sha256=np.array(sha(y)), marked_sha256=np.array(sha(marked)),
        marked_minus_src_rows_0_64=(marked[:64].astype(np.int16) - y[:64]).astype(np.int8),
        bits_clean=np.packbits(bits_clean.astype(np.uint8).reshape(-1)),
        bits_marked=np.packbits(bits_marked.astype(np.uint8).reshape(-1)),
        pattern_marked=DeShuffler(key=KEY).set_shape(PAYLOAD.shape).degenerate(bits_marked))

    # ---------------------------------------------------------------- payload side
    print("payload side: permutations, ties, grayscale")
    pay = {"permutations": {}, "degenerate_cases": [], "grayscale": {}}
    for length, key in ((8, 0), (8, 1), (8, 12345), (16, 0), (5, 7), (441, 0), (1, 0)):
        idx = DeShuffler(key=key).set_shape((length,)).payload_idx
        same(o_pay.permutation(length, key), idx, f"permutation L={length} key={key}")
        pay["permutations"][f"{length}:{key}"] = idx.tolist()
    rng = np.random.RandomState(7)
    for case in range(40):
        length = [8, 8, 8, 5, 16][case % 5]
        n = [32400, 1200, 129600, 97, 4050][case % 5]
        if case % 4 == 0:      # exact ties: every position sees the same count
            base = (np.arange(n) // length) % 2
            bits = base.astype(np.float64)
        elif case % 4 == 1:    # near ties
            bits = (rng.rand(n) < 0.5).astype(np.float64)
        else:                  # a real mark with noise
            wm = Shuffler(key=case).generate_wm(rng.randint(0, 2, length), (1, n))
            flip = rng.rand(n) < 0.3
            bits = np.where(flip, 1 - wm[0], wm[0]).astype(np.float64)
        ref = DeShuffler(key=case).set_shape((length,)).degenerate(bits.reshape(1, -1))
        same(o_pay.degenerate(bits.reshape(1, -1), length, case), ref, f"degenerate case {case}")
        counts = [int(bits[i::length].sum()) for i in range(length)]
        pay["degenerate_cases"].append({"length": length, "n": n, "key": case,
                                        "counts": counts, "pattern": ref.tolist()})
    qr = cv2.imread(os.path.join(REF, "tests", "media", "wms", "qr.jpeg"), cv2.IMREAD_GRAYSCALE)
    wm_ref = GrayScale(key=KEY).generate_wm(qr, (1, 32400))
    same(o_pay.generate_wm_grayscale(qr, (1, 32400), KEY), wm_ref, "grayscale generate_wm")
    back = DeGrayScale(key=KEY).set_shape(qr.shape).degenerate(wm_ref.astype(np.float64))
    same(o_pay.degenerate_grayscale(wm_ref.astype(np.float64), qr.shape, KEY), back, "grayscale degenerate")
    pay["grayscale"] = {"shape": list(qr.shape), "image": qr.tolist(),
                        "wm_sha256": sha(wm_ref.astype(np.uint8)), "roundtrip": back.tolist()}
    with open(os.path.join(OUT, "payload.json"), "w") as f:
        json.dump(pay, f)
    with open(os.path.join(OUT, "META.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
