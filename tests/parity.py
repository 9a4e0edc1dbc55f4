"""Shared helpers for the parity tests (CUDA path vs oracle)."""
import numpy as np

from oracle import dwt_dct_svd as o_svd

PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])     # tests/mark.py:22 in the reference
KEY = 0
# Width of the band around a quantisation boundary inside which the reference's own float32 rounding decides:
# measured |sigma_0(reference: cv2.dct + sgesdd) - sigma_0(float64)| <= 1.6e-7 * sigma_0 on the fixtures, the
# kernels' own bound is 1.2e-7; 2^-21 = 4.8e-7 covers both (the kernels use the same figure, dwtsvd_tile.cuh).
REL_TOL = 2.0 ** -21
MAX_MASKED = 1e-3                                 # no test may hide more than this fraction of the blocks


def flat_tiles(plane):
    """Per walked tile: are all 64 samples equal?  The reference is exact (hence deterministic) on those."""
    nr, nc = o_svd.block_grid(plane.shape[0], plane.shape[1])
    t = np.asarray(plane)[:nr * 8, :nc * 8].reshape(nr, 8, nc, 8).transpose(0, 2, 1, 3).reshape(nr * nc, 64)
    return (t == t[:, :1]).all(axis=1) if t.size else np.zeros(0, dtype=bool)


def knife_edge_blocks(plane_f32, scale=15.0, rel_tol=REL_TOL, check=True):
    """Blocks whose sigma_0 (float64 ground truth on the float32 LL band) lies within ``rel_tol * sigma_0`` of
    a quantisation boundary (k*scale or (k+1/2)*scale).  There the reference's own float32 cv2.dct + LAPACK
    rounding decides the bit / the floor, so no independent implementation can be expected to agree;
    everywhere else - and on FLAT tiles even there, where the reference is exact - agreement must be exact."""
    yuv = np.zeros(plane_f32.shape + (3,), dtype=np.float32)
    yuv[:, :, 1] = plane_f32
    _, s64 = o_svd.decode_sigma(yuv)
    r = np.mod(s64, scale)
    tol = rel_tol * s64
    half = 0.5 * scale
    edge_floor = ((r < tol) | (scale - r < tol)) & (s64 > 0)
    edge_bit = ((np.abs(r - half) < tol) & (s64 > 0)) | edge_floor
    flat = flat_tiles(plane_f32)
    edge_bit &= ~flat
    edge_floor &= ~flat
    if check:
        assert_mask_is_small(edge_bit, f"{plane_f32.shape} plane")
    return edge_bit, edge_floor, s64


def assert_mask_is_small(mask, what=""):
    frac = float(np.mean(mask)) if np.size(mask) else 0.0
    if np.size(mask) < 4096:                      # small planes: one or two blocks are already above any fraction
        assert np.sum(mask) <= 2, f"{what}: {int(np.sum(mask))} of {np.size(mask)} blocks on a quantisation boundary"
        return frac
    assert frac < MAX_MASKED, f"{what}: the knife-edge mask hides {frac:.2e} of the blocks (limit {MAX_MASKED:.0e})"
    return frac


def tile_mask_to_pixels(mask_tiles, shape):
    """Per-tile boolean mask -> per-pixel mask over the walked area of a plane of ``shape``."""
    nr, nc = o_svd.block_grid(shape[0], shape[1])
    m = np.zeros(shape, dtype=bool)
    m[:nr * 8, :nc * 8] = np.kron(mask_tiles.reshape(nr, nc), np.ones((8, 8), dtype=bool))
    return m
