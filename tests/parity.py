"""Shared helpers for the parity tests (CUDA path vs oracle)."""
import numpy as np

from oracle import dwt_dct_svd as o_svd

PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])     # tests/mark.py:22 in the reference
KEY = 0


def knife_edge_blocks(plane_f32, scale=15.0, tol=2e-3):
    """Blocks whose sigma_0 (float64 ground truth on the float32 LL band) lies within ``tol`` of a
    quantisation boundary (k*scale or (k+1/2)*scale).  There the reference's own float32
    cv2.dct + LAPACK rounding decides the bit / the floor, so no independent implementation can
    be expected to agree; everywhere else agreement must be exact."""
    yuv = np.zeros(plane_f32.shape + (3,), dtype=np.float32)
    yuv[:, :, 1] = plane_f32
    _, s64 = o_svd.decode_sigma(yuv)
    r = np.mod(s64, scale)
    half = 0.5 * scale
    edge_bit = (np.abs(r - half) < tol) | (r < tol) | (scale - r < tol)
    edge_floor = (r < tol) | (scale - r < tol)
    return edge_bit, edge_floor, s64


def tile_mask_to_pixels(mask_tiles, shape):
    """Per-tile boolean mask -> per-pixel mask over the walked area of a plane of ``shape``."""
    nr, nc = o_svd.block_grid(shape[0], shape[1])
    m = np.zeros(shape, dtype=bool)
    m[:nr * 8, :nc * 8] = np.kron(mask_tiles.reshape(nr, nc), np.ones((8, 8), dtype=bool))
    return m
