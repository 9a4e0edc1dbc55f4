"""GPU: detection after distortions (BASELINE config 5): the CUDA extractor must track the reference
extractor on the SAME attacked frames - raw-bit agreement away from quantisation boundaries and
identical voted payloads - whatever the attack does to the mark itself."""
import numpy as np
import pytest
import torch

from oracle import attacks, dwt_dct_svd as o_svd, payload as o_pay, synth
from parity import PAYLOAD, KEY, knife_edge_blocks

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W = 1080, 1920

ATTACKS = [
    ("none", lambda p, f: p),
    ("jpeg_q95", lambda p, f: attacks.jpeg_requant(p, 95)),
    ("jpeg_q75", lambda p, f: attacks.jpeg_requant(p, 75)),
    ("noise_s1", lambda p, f: attacks.gaussian_noise(p, 1.0, f)),
    ("noise_s2", lambda p, f: attacks.gaussian_noise(p, 2.0, f)),
    ("noise_s4", lambda p, f: attacks.gaussian_noise(p, 4.0, f)),
    ("resize_720p", lambda p, f: attacks.resize_roundtrip(p)),
]


@pytest.fixture(scope="module")
def marked_frames():
    from b200wm import ops
    frames = np.stack([synth.luma_plane_u8(H, W, f, 555) for f in range(2)])
    wm = o_pay.generate_wm(PAYLOAD, (1, H * W // 64), KEY)
    t = torch.from_numpy(frames).to(DEV)
    packed, n = ops.pack_bits(wm[0], device=DEV)
    ops.dwtsvd_embed_(t, packed, n)
    return t.cpu().numpy(), wm[0]


@pytest.mark.parametrize("name,attack", ATTACKS, ids=[a[0] for a in ATTACKS])
def test_extractor_tracks_reference_under_attack(marked_frames, name, attack):
    from b200wm import ops
    marked, wm = marked_frames
    n = H * W // 64
    for f in range(marked.shape[0]):
        attacked = attack(marked[f], f)
        want = o_svd.extract_plane(attacked)[0]
        raw, counts = ops.dwtsvd_extract(torch.from_numpy(attacked).to(DEV), payload_len=8)
        got = ops.unpack_bits(raw, n)[0]
        diff = np.flatnonzero(got != want)
        edge, _, _ = knife_edge_blocks(attacked.astype(np.float32), check=False)
        assert edge.mean() < (3e-3 if name.startswith("jpeg") else 1e-3), (name, edge.mean())     # requantised planes: see below
        assert all(edge[c] for c in diff), f"{name}: raw bits differ away from quantisation boundaries"
        assert diff.size <= 0.002 * n
        ber_ref = float((want != wm).mean())
        ber_gpu = float((got != wm).mean())
        assert abs(ber_ref - ber_gpu) < 2e-3
        perm = torch.from_numpy(o_pay.permutation(8, KEY).astype(np.int32)).to(DEV)
        patterns, _ = ops.vote_finish(counts, n, perm)
        assert np.array_equal(patterns[0].cpu().numpy(), o_pay.degenerate(want.reshape(1, -1), 8, KEY)), name
        print(f"{name}: raw BER reference {ber_ref:.4f} gpu {ber_gpu:.4f}, boundary flips {diff.size}, "
              f"payload {'ok' if np.array_equal(patterns[0].cpu().numpy(), PAYLOAD) else 'lost'}")


def test_gpu_attack_kernels_match_cpu_definitions(marked_frames):
    """The GPU distortion kernels implement oracle/attacks.py: the noise kernel exactly (same noise
    field), the JPEG-like requantiser up to rounding ties: the DCT of an integer block has
    coefficients on exact multiples of 1/8 (DC = sum/8, and [0][4], [4][0], [4][4] alike), so with a
    quantiser step of 2 or 4 a few percent of the coefficients sit EXACTLY on a rounding tie that float32
    noise inside any DCT implementation breaks either way (cv2.dct vs an exact float64 DCT differ on
    2.7 % of the samples at quality 95, by at most 2 levels)."""
    from b200wm import ops
    marked, _ = marked_frames
    rng = np.random.RandomState(3)
    noise = rng.normal(0, 2.0, marked.shape).astype(np.float32)
    want = np.clip(np.rint(marked.astype(np.float32) + noise), 0, 255).astype(np.uint8)
    t = torch.from_numpy(marked.copy()).to(DEV)
    ops.attack_add_noise_(t, torch.from_numpy(noise).to(DEV))
    assert np.array_equal(t.cpu().numpy(), want)
    from scipy.fft import dctn, idctn
    for q in (95, 75, 50):
        want = attacks.jpeg_requant(marked[0], q)
        t = torch.from_numpy(marked[:1].copy()).to(DEV)
        ops.attack_jpeg_requant_(t, q)
        got = t.cpu().numpy()[0]
        d = np.abs(got.astype(np.int16) - want)
        # Every block that differs must be EXPLAINED by a rounding tie, checked against an exact (float64) DCT:
        #  (a) a coefficient whose c/Q lies within 1e-4 of k + 1/2 (np.rint may go either way on float32 noise), or
        #  (b) identical quantised coefficients, and the differing samples are 1 level apart with an exact value
        #      within 1e-3 of k + 1/2 before the final rounding.
        table = attacks.jpeg_table(q).astype(np.float64)
        bad = np.argwhere(d.reshape(H // 8, 8, W // 8, 8).max(axis=(1, 3)) > 0)
        coeff_ties = round_ties = 0
        for by, bx in bad:
            blk = marked[0, by * 8:by * 8 + 8, bx * 8:bx * 8 + 8].astype(np.float64) - 128.0
            ratio = dctn(blk, norm="ortho") / table
            if (0.5 - np.abs(ratio - np.rint(ratio))).min() < 1e-4:
                coeff_ties += 1
                continue
            exact = idctn(np.rint(ratio) * table, norm="ortho") + 128.0
            dd = d[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8]
            assert dd.max() == 1 and (np.abs(np.abs(exact - np.floor(exact)) - 0.5)[dd > 0] < 1e-3).all(), (q, by, bx)
            round_ties += 1
        print(f"jpeg q{q}: {len(bad)} of {H * W // 64} blocks differ: {coeff_ties} coefficient ties, {round_ties} final-rounding ties")
        assert len(bad) < 0.08 * (H * W // 64)
        # and the two extractors agree on what is left of the mark after the GPU attack; requantised planes sit on
        # quantisation boundaries more often than camera content (sums of few lattice coefficients), hence the limit
        raw, _ = ops.dwtsvd_extract(t)
        bits = ops.unpack_bits(raw, H * W // 64)[0]
        ref_bits = o_svd.extract_plane(got)[0]
        edge, _, _ = knife_edge_blocks(got.astype(np.float32), check=False)
        assert edge.mean() < 3e-3, edge.mean()
        assert all(edge[c] for c in np.flatnonzero(bits != ref_bits))


@pytest.mark.parametrize("src_hw,dst_hw", [((1080, 1920), (720, 1280)), ((240, 320), (160, 213)), ((64, 96), (48, 40)),
                                           ((37, 53), (20, 31)), ((48, 96), (24, 48)), ((48, 96), (16, 32)),
                                           ((48, 96), (24, 32)), ((30, 50), (30, 25)), ((9, 11), (1, 1))])
def test_gpu_resize_is_cv2_bit_for_bit(src_hw, dst_hw):
    """b200wm_attack_resize == cv2.resize on uint8 planes, INTER_AREA down and INTER_LINEAR both ways,
    batched, with a pitched (non-contiguous) source."""
    import cv2
    from b200wm import ops
    rng = np.random.RandomState(src_hw[1] + dst_hw[0])
    frames = rng.randint(0, 256, (3,) + src_hw + (2,)).astype(np.uint8)
    src = frames[..., 0]
    dsize = (dst_hw[1], dst_hw[0])
    t = torch.from_numpy(frames).to(DEV)[..., 0].permute(0, 1, 2)       # element stride 2 is not planar:
    t = t.contiguous()[:, :, :]                                          # planar copy ...
    pad = torch.zeros((3, src_hw[0], src_hw[1] + 5), dtype=torch.uint8, device=DEV)
    pad[:, :, :src_hw[1]] = t                                            # ... inside wider rows (pitch != width)
    view = pad[:, :, :src_hw[1]]
    small = ops.attack_resize(view, dsize, ops.INTER_AREA)
    for f in range(3):
        assert np.array_equal(small[f].cpu().numpy(), cv2.resize(src[f], dsize, interpolation=cv2.INTER_AREA)), f
    back = ops.attack_resize(small, (src_hw[1], src_hw[0]), ops.INTER_LINEAR)
    down = ops.attack_resize(view, dsize, ops.INTER_LINEAR)
    for f in range(3):
        want_small = cv2.resize(src[f], dsize, interpolation=cv2.INTER_AREA)
        assert np.array_equal(back[f].cpu().numpy(), cv2.resize(want_small, (src_hw[1], src_hw[0]), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(down[f].cpu().numpy(), cv2.resize(src[f], dsize, interpolation=cv2.INTER_LINEAR))


def test_gpu_resize_roundtrip_and_errors(marked_frames):
    from b200wm import ops
    marked, _ = marked_frames
    t = torch.from_numpy(marked.copy()).to(DEV)
    ops.attack_resize_roundtrip_(t)
    for f in range(marked.shape[0]):
        assert np.array_equal(t[f].cpu().numpy(), attacks.resize_roundtrip(marked[f]))
    with pytest.raises(Exception):        # INTER_AREA enlargement is a different filter in OpenCV: refused
        ops.attack_resize(t, (W * 2, H * 2), ops.INTER_AREA)
    with pytest.raises(Exception):
        ops.attack_resize(t, (W // 2, H // 2), 2)    # INTER_CUBIC


def test_gpu_resize_random_geometries():
    """Seeded sweep over odd sizes and ratios: b200wm_attack_resize stays bit-identical to cv2.resize."""
    import cv2
    from b200wm import ops
    rng = np.random.RandomState(2026)
    for _ in range(40):
        sh, sw = int(rng.randint(2, 90)), int(rng.randint(2, 120))
        dh, dw = int(rng.randint(1, sh + 1)), int(rng.randint(1, sw + 1))
        src = rng.randint(0, 256, (sh, sw)).astype(np.uint8)
        t = torch.from_numpy(src).to(DEV)
        got = ops.attack_resize(t, (dw, dh), ops.INTER_AREA).cpu().numpy()
        assert np.array_equal(got, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)), ((sh, sw), (dh, dw))
        uh, uw = int(rng.randint(1, 140)), int(rng.randint(1, 160))
        got = ops.attack_resize(t, (uw, uh), ops.INTER_LINEAR).cpu().numpy()
        assert np.array_equal(got, cv2.resize(src, (uw, uh), interpolation=cv2.INTER_LINEAR)), ((sh, sw), (uh, uw))


def test_jpeg_requant_vector_and_generic_paths_agree():
    """The conversion-free 64-bit path of the JPEG-like requantiser (8-byte aligned rows) and the byte-wise generic path
    (unaligned view) give the same planes, in place and batched; nothing outside the 8x8-covered area is written."""
    from b200wm import ops
    rng = np.random.RandomState(12)
    base = rng.randint(0, 256, (3, 75, 203)).astype(np.uint8)
    base[0, :16, :24] = 255
    base[1, 8:24, 40:64] = 0
    for q in (95, 60, 20):
        a = torch.from_numpy(base.copy()).to(DEV)
        aligned = torch.zeros((3, 75, 208), dtype=torch.uint8, device=DEV)[:, :, :203]      # pitch 208: vector path
        aligned.copy_(a)
        big = torch.zeros((3, 77, 211), dtype=torch.uint8, device=DEV)
        view = big[:, 1:76, 5:208]                                                              # pitch 211, offset 5: generic path
        view.copy_(a)
        ops.attack_jpeg_requant_(aligned, q)
        ops.attack_jpeg_requant_(view, q)
        assert torch.equal(aligned, view), q
        got = aligned.cpu().numpy()
        assert np.array_equal(got[:, 72:], base[:, 72:]) and np.array_equal(got[:, :, 200:], base[:, :, 200:])
        want = attacks.jpeg_requant(base[2], q)
        assert (np.abs(got[2].astype(np.int16) - want) > 0).mean() < 0.06
