"""Seeded synthetic inputs shared by the golden generator and the parity tests.

Test infrastructure (see oracle/__init__.py).  The reference ships one 240p
clip and one 1080p still and nothing else, so larger and
odd-sized inputs are generated.  ``luma_plane_u8`` follows the recipe of
SURVEY.md §8(d) config 2 (smooth moving pattern + Gaussian noise, video range).
"""
import numpy as np


def random_bgr(h, w, seed):
    """Smooth gradients plus noise, uint8 H x W x 3 (a stand-in for a decoded frame)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([128 + 90 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], axis=2)
    return np.clip(np.round(base + rng.normal(0, 12, (h, w, 3))), 0, 255).astype(np.uint8)


def luma_plane_u8(h, w, frame_index, seed, sigma=6.0):
    rng = np.random.RandomState((seed + frame_index) % (2 ** 32))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    f = float(frame_index)
    pattern = 128 + 80 * np.sin(2 * np.pi * (3 * xx / w + f / 97)) * np.cos(2 * np.pi * (2 * yy / h + f / 53))
    return np.clip(np.round(pattern + rng.normal(0, sigma, (h, w))), 16, 235).astype(np.uint8)


def full_range_plane_u8(h, w, seed):
    """Uniform full-range noise with saturated patches: exercises clipping at 0/255,
    all-zero blocks and blocks whose top two singular values are close."""
    rng = np.random.RandomState(seed)
    p = rng.randint(0, 256, (h, w)).astype(np.uint8)
    if h >= 32 and w >= 32:
        p[: h // 4, : w // 4] = 0
        p[h // 4: h // 2, : w // 4] = 255
        p[: h // 4, w // 4: w // 2] = rng.randint(0, 4, (h // 4, w // 2 - w // 4))
        p[h // 2: 3 * h // 4, : w // 4] = rng.randint(252, 256, (3 * h // 4 - h // 2, w // 4))
    return p
