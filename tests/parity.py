"""Shared helpers for the parity tests (CUDA path vs oracle)."""
import numpy as np

from oracle import knife_edge
from oracle.knife_edge import REL_TOL, flat_tiles, tile_mask_to_pixels     # noqa: F401  (re-exported)

PAYLOAD = np.array([0, 1, 1, 0, 0, 1, 0, 1])     # tests/mark.py:22 in the reference
KEY = 0
MAX_MASKED = 1e-3                                 # no test may hide more than this fraction of the blocks


def knife_edge_blocks(plane_f32, scale=15.0, rel_tol=REL_TOL, check=True):
    """oracle.knife_edge.knife_edge_blocks, plus the rule that a test may not hide more than MAX_MASKED of its
    blocks behind the mask (``check=False`` for callers that state their own limit)."""
    edge_bit, edge_floor, s64 = knife_edge.knife_edge_blocks(plane_f32, scale, rel_tol)
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
