"""Distortions for the robustness study (BASELINE.json config 5).

Test infrastructure (see oracle/__init__.py).  The reference has NO attack module: the only
distortions it ever applies are a JPEG round trip (tests/test.py:99,111) and x264/HLS transcodes
through ffmpeg (src/offmark/video/frame_writer.py:31-37).  SURVEY.md §8d config 5 therefore
defines synthetic stand-ins, restated here on the CPU; the same definitions are implemented on the
GPU in video-fingerprinting_b200/csrc/attacks.cu and compared with these in tests/.
"""
import math

import numpy as np
import cv2

# libjpeg luminance quantisation table (Annex K of ITU-T T.81)
JPEG_LUMA = np.array([
    [16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55],
    [14, 13, 16, 24, 40, 57, 69, 56], [14, 17, 22, 29, 51, 87, 80, 62],
    [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
    [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.float32)


def jpeg_table(quality):
    """libjpeg's quality scaling of the luminance table (jcparam.c jpeg_quality_scaling)."""
    q = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((JPEG_LUMA * q + 50) // 100, 1, 255).astype(np.float32)


def jpeg_requant(plane_u8, quality):
    """8x8 DCT of (plane - 128), quantise with the scaled table, dequantise, IDCT, round, clip."""
    h, w = plane_u8.shape
    h8, w8 = h // 8 * 8, w // 8 * 8
    t = jpeg_table(quality)
    out = plane_u8.copy()
    f = plane_u8[:h8, :w8].astype(np.float32) - 128.0
    blocks = f.reshape(h8 // 8, 8, w8 // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 8, 8)
    res = np.empty_like(blocks)
    for i, b in enumerate(blocks):
        c = cv2.dct(b)
        res[i] = cv2.idct(np.rint(c / t) * t)
    res = res.reshape(h8 // 8, w8 // 8, 8, 8).transpose(0, 2, 1, 3).reshape(h8, w8)
    out[:h8, :w8] = np.clip(np.rint(res + 128.0), 0, 255).astype(np.uint8)
    return out


def gaussian_noise(plane_u8, sigma, seed):
    rng = np.random.RandomState(seed)
    return np.clip(np.rint(plane_u8.astype(np.float32) + rng.normal(0, sigma, plane_u8.shape)), 0, 255).astype(np.uint8)


def resize_roundtrip(plane_u8, scale=2.0 / 3.0):
    """Area down-scale then bilinear up-scale back (1080p -> 720p -> 1080p for scale 2/3)."""
    h, w = plane_u8.shape
    small = cv2.resize(plane_u8, (int(round(w * scale)), int(round(h * scale))), interpolation=cv2.INTER_AREA)
    return cv2.resize(small, (w, h), interpolation=cv2.INTER_LINEAR)


# ---- restatement of cv2.resize on uint8 (what csrc/attacks.cu implements) ---------------------------
# resize_roundtrip() above calls OpenCV itself and is the definition; the functions below restate
# OpenCV's arithmetic step by step (modules/imgproc/src/resize.cpp of OpenCV 4.x: computeResizeAreaTab /
# ResizeArea_Invoker, ResizeAreaFast_Invoker, HResizeLinear / VResizeLinear with INTER_RESIZE_COEF_BITS = 11)
# and are pinned to cv2 bit for bit in tests/test_oracle_golden.py, so that the kernel's comments can be
# checked against something readable.
def _area_table(ssize, dsize):
    scale = ssize / dsize
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cell)))
        for sx in range(s1, s2):
            tab.append((d, sx, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
    return tab


def resize_area_restated(plane_u8, dsize):
    """cv2.resize(plane, dsize=(width, height), interpolation=cv2.INTER_AREA) for reductions."""
    sh, sw = plane_u8.shape
    dw, dh = dsize
    if sw % dw == 0 and sh % dh == 0:
        ix, iy = sw // dw, sh // dh
        s = plane_u8.reshape(dh, iy, dw, ix).astype(np.int64).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(s.astype(np.float32) * (np.float32(1.0) / np.float32(ix * iy))), 0, 255).astype(np.uint8)
    src = plane_u8.astype(np.float32)
    buf = np.zeros((sh, dw), np.float32)
    for d, sx, a in _area_table(sw, dw):
        buf[:, d] = buf[:, d] + src[:, sx] * a
    out = np.zeros((dh, dw), np.float32)
    seen = np.zeros(dh, bool)
    for d, sy, b in _area_table(sh, dh):
        out[d] = out[d] + b * buf[sy] if seen[d] else b * buf[sy]
        seen[d] = True
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def _linear_table(ssize, dsize, clamp_coordinate):
    scale = ssize / dsize
    ofs = np.zeros(dsize, np.int64)
    coef = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_coordinate:                      # columns: xofs / ialpha of cv::resize
            if s < 0:
                f, s = np.float32(0), 0
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        ofs[d] = s
        coef[d, 0] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
        coef[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return ofs, coef


def resize_linear_restated(plane_u8, dsize):
    """cv2.resize(plane, dsize=(width, height), interpolation=cv2.INTER_LINEAR) on uint8."""
    sh, sw = plane_u8.shape
    dw, dh = dsize
    if sw == 2 * dw and sh == 2 * dh:             # cv2 switches an exact 2x2 reduction to INTER_AREA
        return resize_area_restated(plane_u8, dsize)
    xo, xa = _linear_table(sw, dw, True)
    yo, ya = _linear_table(sh, dh, False)         # rows keep their weights; the row index is clamped instead
    src = plane_u8.astype(np.int64)
    rows = src[:, xo] * xa[:, 0] + src[:, np.minimum(xo + 1, sw - 1)] * xa[:, 1]
    r0, r1 = rows[np.clip(yo, 0, sh - 1)], rows[np.clip(yo + 1, 0, sh - 1)]
    v = (((ya[:, 0:1] * (r0 >> 4)) >> 16) + ((ya[:, 1:2] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)
