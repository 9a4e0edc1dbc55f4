"""Distortions for the robustness study (BASELINE.json config 5).

Test infrastructure (see oracle/__init__.py).  The reference has NO attack module: the only
distortions it ever applies are a JPEG round trip (tests/test.py:99,111) and x264/HLS transcodes
through ffmpeg (src/offmark/video/frame_writer.py:31-37).  SURVEY.md §8d config 5 therefore
defines synthetic stand-ins, restated here on the CPU; the same definitions are implemented on the
GPU in video-fingerprinting_b200/csrc/attacks.cu and compared with these in tests/.
"""
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
