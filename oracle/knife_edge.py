"""Which blocks can two float32 implementations legitimately disagree on?

Test infrastructure (see oracle/__init__.py): used by the parity tests and by the parity block of ``bench.py``.

The reference decides ``floor(sigma_0 / scale)`` (embed/dwt_dct_svd_encoder.py:44) and ``sigma_0 % scale > scale/2``
(extract/dwt_dct_svd_decoder.py:36) on a float32 sigma_0 that went through cv2.dct and LAPACK sgesdd; measured on
the fixtures that value is within 1.6e-7 (relative) of the float64 truth, the CUDA kernels are within 1.2e-7.  A
block whose sigma_0 lies closer than 2^-21 (4.8e-7, relative) to a quantisation boundary is therefore decided by
rounding noise inside the reference - a "knife-edge" block - and is excluded from exact comparisons; every other
block, and every FLAT block even on a boundary (the reference is exact there), must agree exactly.
"""
import numpy as np

from . import dwt_dct_svd as o_svd

REL_TOL = 2.0 ** -21


def flat_tiles(plane):
    """Per walked tile: are all 64 samples equal?"""
    nr, nc = o_svd.block_grid(plane.shape[0], plane.shape[1])
    t = np.asarray(plane)[:nr * 8, :nc * 8].reshape(nr, 8, nc, 8).transpose(0, 2, 1, 3).reshape(nr * nc, 64)
    return (t == t[:, :1]).all(axis=1) if t.size else np.zeros(0, dtype=bool)


def knife_edge_blocks(plane_f32, scale=15.0, rel_tol=REL_TOL):
    """-> (edge_bit, edge_floor, sigma64) per walked block: on a bit-or-floor boundary, on a floor boundary, and
    the float64 sigma_0 of the float32 Haar band."""
    yuv = np.zeros(plane_f32.shape + (3,), dtype=np.float32)
    yuv[:, :, 1] = plane_f32
    _, s64 = o_svd.decode_sigma(yuv)
    r = np.mod(s64, scale)
    tol = rel_tol * s64
    edge_floor = ((r < tol) | (scale - r < tol)) & (s64 > 0)
    edge_bit = ((np.abs(r - 0.5 * scale) < tol) & (s64 > 0)) | edge_floor
    flat = flat_tiles(plane_f32)
    return edge_bit & ~flat, edge_floor & ~flat, s64


def tile_mask_to_pixels(mask_tiles, shape):
    """Per-tile boolean mask -> per-pixel mask over the walked area of a plane of ``shape``."""
    nr, nc = o_svd.block_grid(shape[0], shape[1])
    m = np.zeros(shape, dtype=bool)
    m[:nr * 8, :nc * 8] = np.kron(mask_tiles.reshape(nr, nc), np.ones((8, 8), dtype=bool))
    return m
