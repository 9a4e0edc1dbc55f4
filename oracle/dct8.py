"""CPU restatement of the 8x8 block-DCT quantisation-index pair.

Test infrastructure (see oracle/__init__.py).  Follows
src/offmark/embed/dct_encoder.py:18-102 (embed + masks) and
src/offmark/extract/dct_decoder.py:10-89 (extract; the masks there are a
verbatim duplicate of the encoder's).

The per-block ``cv2.dct`` calls are kept (OpenCV's 8x8 2-D transform is not
bit-reproducible from 1-D passes), everything after them is vectorised with
the reference's float64 expressions written in the same order.
"""
import numpy as np
import cv2

BLK = 8


def _blocks8(plane):
    rows, cols = plane.shape[0] // BLK, plane.shape[1] // BLK
    return rows, cols


def _dct_all(plane):
    """float32 (rows, cols, 8, 8): ``cv2.dct`` of every full 8x8 block."""
    rows, cols = _blocks8(plane)
    out = np.empty((rows, cols, BLK, BLK), dtype=np.float32)
    src = np.asarray(plane, dtype=np.float32)
    for i in range(rows):
        for j in range(cols):
            out[i, j] = cv2.dct(np.ascontiguousarray(src[i * BLK:(i + 1) * BLK, j * BLK:(j + 1) * BLK]))
    return out


def luminance_mask_from_dc(dc):
    """dct_encoder.py:52-67 given the float32 DC coefficient of each Y block."""
    l_min, l_max, f_max = 90, 255, 2
    mask = dc.astype(np.float64)          # stored into an np.zeros float64 array (:45,51)
    mask /= 8
    mean = max(l_min, np.mean(mask))
    f_ref = 1 + (mean - l_min) * (f_max - 1) / (l_max - l_min)
    with np.errstate(divide='ignore', invalid='ignore'):
        bright = 1 + (mask - mean) / (l_max - mean) * (f_max - f_ref)
    return np.where(mask > mean, bright,
                    np.where(mask < 15, 1.25, np.where(mask < 25, 1.125, 1.0)))


def texture_mask_terms(coeffs):
    """The quantities the decision tree of dct_encoder.py:80-101 compares, float32 like the reference's: sums
    accumulated left to right as the Python expressions do, ``np.sum`` over the 64 magnitudes numpy's own."""
    c = np.abs(coeffs)
    dcl = c[..., 0, 0] + c[..., 0, 1] + c[..., 0, 2] + c[..., 1, 0] + c[..., 1, 1] + c[..., 2, 0]
    total = np.array([[np.sum(c[i, j]) for j in range(c.shape[1])] for i in range(c.shape[0])],
                     dtype=np.float32).reshape(c.shape[:2])
    eh = total - dcl
    e = (c[..., 3, 0] + c[..., 4, 0] + c[..., 5, 0] + c[..., 6, 0]
         + c[..., 0, 3] + c[..., 0, 4] + c[..., 0, 5] + c[..., 0, 6]
         + c[..., 2, 1] + c[..., 1, 2] + c[..., 2, 2] + c[..., 3, 3])
    h = eh - e
    l = dcl - c[..., 0, 0]
    with np.errstate(divide='ignore', invalid='ignore'):
        l_e = l / e
        le_h = (l + e) / h
    return {"eh": eh, "e": e, "h": h, "l": l, "l_e": l_e, "le_h": le_h}


def texture_mask_from_coeffs(coeffs):
    """dct_encoder.py:80-101 given ``cv2.dct`` of each Y block, float32 (rows, cols, 8, 8)."""
    t = texture_mask_terms(coeffs)
    eh, e, h, l, l_e, le_h = t["eh"], t["e"], t["h"], t["l"], t["l_e"], t["le_h"]
    a1, b1, a2, b2 = 2.3, 1.6, 1.4, 1.1
    ramp = 1 + 1.25 * (eh - np.float32(290)) / (1800 - 290)
    ramp = ramp.astype(np.float64)
    flat_or_strong = np.where(l + e <= 400, 1.125, 1.25)
    hi_rule = ((l_e >= a2) & (le_h >= b2)) | ((l_e >= b2) & (le_h >= a2)) | (le_h > 4)
    lo_rule = ((l_e >= a1) & (le_h >= b1)) | ((l_e >= b1) & (le_h >= a1)) | (le_h > 4)
    hi_val = np.where(hi_rule, flat_or_strong, ramp)
    lo_val = np.where(lo_rule, flat_or_strong, np.where(e + h > 290, ramp, 1.0))
    return np.where(eh > 125, np.where(eh > 900, hi_val, lo_val), 1.0)


def tie_blocks(yuv, alpha=20, abs_tol=5e-3, rel_tol=2e-5):
    """Blocks of an H x W x 3 float32 frame on which two float32 implementations of the 8x8 pair may legitimately
    differ, each for a nameable reason (all tested against the reference's own float32 intermediates):

    ``mask``   a comparison of the mask decision trees is a tie: eh against 125 / 900, e+h against 290, l+e against
               400 (``abs_tol``), a ratio against 2.3 / 1.6 / 1.4 / 1.1 / 4 (``rel_tol``), the block mean against the
               frame mean, 15 or 25 (dct_encoder.py:58-65, :82-101);
    ``sign``   |c21| of the chroma block is non-zero but below float32 noise (< 1e-3), so ``np.sign(c21)``
               (dct_encoder.py:33,35) is noise; an EXACT zero (flat or mirror-symmetric block) is not a tie: both
               implementations must leave such a block unmarked;
    ``floor``  |c21| / (2 step) is within ``rel_tol`` of an integer (the floor of :32,34 may go either way);
    ``round``  c21 / step is within ``rel_tol`` of k + 1/2 (``np.around`` of dct_decoder.py:24 may go either way).
    """
    lum = _dct_all(yuv[:, :, 0])
    t = texture_mask_terms(lum)
    near = lambda x, c, tol: np.abs(x - c) <= tol                                   # noqa: E731
    with np.errstate(invalid='ignore'):
        mask = near(t["eh"], 125, abs_tol) | near(t["eh"], 900, abs_tol) | near(t["e"] + t["h"], 290, abs_tol) | \
            near(t["l"] + t["e"], 400, abs_tol)
        for ratio in (t["l_e"], t["le_h"]):
            for c in (2.3, 1.6, 1.4, 1.1, 4.0):
                mask |= np.isfinite(ratio) & near(ratio, c, rel_tol * c + 1e-6) & (t["eh"] > 125 - abs_tol)
    mean = lum[..., 0, 0].astype(np.float64) / 8
    frame_mean = max(90, np.mean(mean))
    mask |= near(mean, frame_mean, 1e-4) | near(mean, 15, 1e-4) | near(mean, 25, 1e-4)
    step = alpha * (texture_mask_from_coeffs(lum) * luminance_mask_from_dc(lum[..., 0, 0]))
    c21 = _dct_all(yuv[:, :, 1])[..., 2, 1].astype(np.float64)
    q2 = np.abs(c21) / (2 * step)
    q1 = c21 / step
    return {"mask": mask, "sign": (np.abs(c21) < 1e-3) & (c21 != 0),
            "floor": np.abs(q2 - np.rint(q2)) <= rel_tol * np.maximum(q2, 1.0) + 1e-6,
            "round": np.abs(np.abs(q1 - np.floor(q1)) - 0.5) <= rel_tol * np.maximum(np.abs(q1), 1.0) + 1e-6}


def masks(lum):
    """``tex_mask * lum_mask`` (dct_encoder.py:21-23), float64 (rows, cols)."""
    coeffs = _dct_all(lum)
    return texture_mask_from_coeffs(coeffs) * luminance_mask_from_dc(coeffs[..., 0, 0])


def encode(yuv, wm, alpha=20):
    """``DctEncoder.encode`` (dct_encoder.py:18-39): QIM of coefficient [2][1] of
    every 8x8 block of channel 1 with step ``alpha * mask``; mutates ``yuv``."""
    bits = np.asarray(wm[0])
    mask = masks(yuv[:, :, 0])
    rows, cols = mask.shape
    if len(bits) < rows * cols:
        raise IndexError("watermark shorter than the number of blocks")
    chan = yuv[:, :, 1]
    c = 0
    for i in range(rows):
        for j in range(cols):
            win = (slice(i * BLK, (i + 1) * BLK), slice(j * BLK, (j + 1) * BLK))
            coeffs = cv2.dct(np.ascontiguousarray(chan[win]))
            step = alpha * mask[i, j]
            step2 = step + step
            base = np.floor(abs(coeffs[2][1]) / step2) * step2
            if bits[c] != 0:
                base = base + step
            coeffs[2][1] = np.sign(coeffs[2][1]) * base      # sign(0) == 0 kept (:33,35)
            chan[win] = cv2.idct(coeffs)
            c += 1
    return yuv


def decode(yuv, alpha=20):
    """``DctDecoder.decode`` (dct_decoder.py:10-27) -> float64 (1, rows*cols//64)."""
    mask = masks(yuv[:, :, 0])
    rows, cols = mask.shape
    wm = np.zeros(yuv.shape[0] * yuv.shape[1] // BLK // BLK)
    c21 = _dct_all(yuv[:, :, 1])[..., 2, 1].astype(np.float64)   # float32 / float64 step -> float64
    step = alpha * mask
    wm[:rows * cols] = (np.around(c21 / step) % 2 == 1).astype(np.float64).reshape(-1)
    return wm.reshape(1, -1)
