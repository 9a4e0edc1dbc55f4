"""CPU restatement of the Haar-DWT -> block-DCT -> SVD quantisation-index pair.

Test infrastructure (see oracle/__init__.py).  Follows
src/offmark/embed/dwt_dct_svd_encoder.py:19-45 (embed) and
src/offmark/extract/dwt_dct_svd_decoder.py:12-37 (extract).

Two forms are kept:

* ``*_vectorised`` - all blocks of a plane at once with numpy/cv2.  Checked
  bit for bit against the reference itself by oracle/make_golden.py (batched
  ``np.linalg.svd`` runs the same LAPACK ``sgesdd`` per matrix, ``cv2.dct``
  with ``DCT_ROWS`` applied twice reproduces the 4x4 2-D ``cv2.dct`` exactly,
  and batched float32 ``@`` equals ``np.dot`` on 4x4 blocks).  This is the
  checker the parity tests use.
* ``*_per_block`` - one block at a time, the way the reference walks the LL
  band.  Pure-Python loops: small cases, and the CPU baseline that
  ``bench.py`` times because it has the reference's cost structure.
"""
import numpy as np
import cv2

from .haar import dwt2_haar, idwt2_haar

DEFAULT_SCALES = (0, 15, 0)
DEFAULT_BLK = 4


def wm_capacity(frame_shape):
    """embed/dwt_dct_svd_encoder.py:14-17 -> ``(1, rows*cols//64)``."""
    rows, cols = frame_shape[0], frame_shape[1]
    return (1, rows * cols // 64)


def block_grid(rows, cols, blk=DEFAULT_BLK):
    """Blocks walked by the reference for a rows x cols plane: the plane is cut
    to multiples of 4 (encoder.py:24), its LL band is half that, and the walk
    covers ``LL_rows//blk`` x ``LL_cols//blk`` blocks (encoder.py:33-34)."""
    ll_rows, ll_cols = (rows // 4 * 4) // 2, (cols // 4 * 4) // 2
    return ll_rows // blk, ll_cols // blk


# --------------------------------------------------------------------------
# block plumbing
# --------------------------------------------------------------------------
def _to_blocks(ll, blk):
    nr, nc = ll.shape[0] // blk, ll.shape[1] // blk
    return (ll[:nr * blk, :nc * blk].reshape(nr, blk, nc, blk)
            .transpose(0, 2, 1, 3).reshape(nr * nc, blk, blk).copy()), nr, nc


def _from_blocks(ll, blocks, nr, nc, blk):
    ll[:nr * blk, :nc * blk] = (blocks.reshape(nr, nc, blk, blk)
                                .transpose(0, 2, 1, 3).reshape(nr * blk, nc * blk))


def _dct_blocks(blocks, inverse=False):
    """2-D orthonormal DCT-II (``cv2.dct``) of every blk x blk block."""
    n, blk, _ = blocks.shape
    fn = cv2.idct if inverse else cv2.dct
    if blk == 4:
        # rows pass, transpose, rows pass: bit-identical to the 2-D call for 4x4
        t = fn(np.ascontiguousarray(blocks).reshape(n * blk, blk), flags=cv2.DCT_ROWS)
        t = np.ascontiguousarray(t.reshape(n, blk, blk).transpose(0, 2, 1))
        t = fn(t.reshape(n * blk, blk), flags=cv2.DCT_ROWS)
        return np.ascontiguousarray(t.reshape(n, blk, blk).transpose(0, 2, 1))
    return np.stack([fn(b) for b in blocks]) if n else blocks.copy()


# --------------------------------------------------------------------------
# vectorised form (the checker)
# --------------------------------------------------------------------------
def embed_ll_vectorised(ll, wm, scale, blk=DEFAULT_BLK):
    """encoder.py:29-45 on one LL band, in place.  ``wm`` is the 1-D bit array
    (``wm[0]`` of what ``read_wm`` receives)."""
    blocks, nr, nc = _to_blocks(ll, blk)
    n = nr * nc
    if n == 0:
        return ll
    if len(wm) < n:
        raise IndexError("watermark shorter than the number of blocks")  # encoder.py:36
    bits = np.asarray(wm[:n])
    u, s, v = np.linalg.svd(_dct_blocks(blocks))
    # s[0] = (s[0] // scale + 0.25 + 0.5 * wm_bit) * scale, stored back as float32
    s0 = (s[:, 0] // scale + 0.25 + 0.5 * bits) * scale
    s[:, 0] = s0.astype(np.float32)
    rebuilt = u @ (s[:, :, None] * v)
    _from_blocks(ll, _dct_blocks(rebuilt.astype(np.float32), inverse=True), nr, nc, blk)
    return ll


def extract_ll_vectorised(ll, scale, blk=DEFAULT_BLK):
    """decoder.py:23-37 on one LL band -> 0/1 per block (float64), plus sigma_0."""
    blocks, nr, nc = _to_blocks(ll, blk)
    if nr * nc == 0:
        return np.zeros(0), np.zeros(0, dtype=np.float32)
    s = np.linalg.svd(_dct_blocks(blocks), compute_uv=False)
    bits = ((s[:, 0] % scale) > scale * 0.5).astype(np.float64)
    return bits, s[:, 0]


def encode(yuv, wm, scales=DEFAULT_SCALES, blk=DEFAULT_BLK):
    """``DwtDctSvdEncoder.encode`` (encoder.py:19-27): mutates and returns ``yuv``
    (float32 H x W x 3).  ``wm`` is the 2-D array handed to ``read_wm``."""
    rows, cols, _ = yuv.shape
    r4, c4 = rows // 4 * 4, cols // 4 * 4
    for ch in range(3):
        if scales[ch] <= 0:
            continue
        ca, hvd = dwt2_haar(yuv[:r4, :c4, ch])
        embed_ll_vectorised(ca, wm[0], scales[ch], blk)
        yuv[:r4, :c4, ch] = idwt2_haar((ca, hvd))
    return yuv


def decode(yuv, scales=DEFAULT_SCALES, blk=DEFAULT_BLK):
    """``DwtDctSvdDecoder.decode`` (decoder.py:12-21) -> float64 (1, rows*cols//4//blk**2).
    Row 1 is returned whatever ``scales`` says, as in the reference (decoder.py:21)."""
    rows, cols, _ = yuv.shape
    block_num = rows * cols // 4 // (blk * blk)
    wm_bits = np.zeros((3, block_num))
    r4, c4 = rows // 4 * 4, cols // 4 * 4
    for ch in range(3):
        if scales[ch] <= 0:
            continue
        ca, _ = dwt2_haar(yuv[:r4, :c4, ch])
        bits, _ = extract_ll_vectorised(ca, scales[ch], blk)
        wm_bits[ch, :len(bits)] = bits
    return wm_bits[1].reshape(1, -1)


def decode_sigma(yuv, channel=1, blk=DEFAULT_BLK):
    """sigma_0 per block as the reference's float32 LAPACK path sees it, and as
    float64 ground truth on the same float32 LL band (no DCT: it is orthogonal).
    Used by parity tests to tell threshold-noise flips from real mismatches."""
    rows, cols, _ = yuv.shape
    ca, _ = dwt2_haar(yuv[:rows // 4 * 4, :cols // 4 * 4, channel])
    blocks, _, _ = _to_blocks(ca, blk)
    s32 = np.linalg.svd(_dct_blocks(blocks), compute_uv=False)[:, 0]
    s64 = np.linalg.svd(blocks.astype(np.float64), compute_uv=False)[:, 0]
    return s32, s64


# --------------------------------------------------------------------------
# planar uint8 mode (north_star benchmark layout): the marked plane is an
# H x W uint8 luma plane; float32 in, clip/around/uint8 out exactly as the
# caller bracket does (src/offmark/video/embedder.py:37-38).
# --------------------------------------------------------------------------
def embed_plane_u8(plane_u8, wm_bits, scale=15, blk=DEFAULT_BLK):
    rows, cols = plane_u8.shape
    r4, c4 = rows // 4 * 4, cols // 4 * 4
    f = plane_u8.astype(np.float32)
    ca, hvd = dwt2_haar(f[:r4, :c4])
    embed_ll_vectorised(ca, wm_bits, scale, blk)
    f[:r4, :c4] = idwt2_haar((ca, hvd))
    return np.around(np.clip(f, 0, 255)).astype(np.uint8)


def extract_plane(plane, scale=15, blk=DEFAULT_BLK):
    """Bits of one plane (uint8 or float32), padded with zeros to rows*cols//64
    exactly like decoder.py:14-15 does."""
    rows, cols = plane.shape
    block_num = rows * cols // 4 // (blk * blk)
    ca, _ = dwt2_haar(np.asarray(plane[:rows // 4 * 4, :cols // 4 * 4], dtype=np.float32))
    bits, _ = extract_ll_vectorised(ca, scale, blk)
    out = np.zeros(block_num)
    out[:len(bits)] = bits
    return out.reshape(1, -1)


# --------------------------------------------------------------------------
# per-block form (reference cost structure; small cases and CPU baseline)
# --------------------------------------------------------------------------
def _embed_one(block, bit, scale):
    u, s, v = np.linalg.svd(cv2.dct(block))
    s[0] = (s[0] // scale + 0.25 + 0.5 * bit) * scale
    return cv2.idct(np.dot(u, np.dot(np.diag(s), v)))


def _extract_one(block, scale):
    s = np.linalg.svd(cv2.dct(block))[1]
    return int((s[0] % scale) > scale * 0.5)


def encode_per_block(yuv, wm, scales=DEFAULT_SCALES, blk=DEFAULT_BLK):
    rows, cols, _ = yuv.shape
    r4, c4 = rows // 4 * 4, cols // 4 * 4
    bits = wm[0]
    for ch in range(3):
        scale = scales[ch]
        if scale <= 0:
            continue
        ca, hvd = dwt2_haar(yuv[:r4, :c4, ch])
        nr, nc = ca.shape[0] // blk, ca.shape[1] // blk
        for c in range(nr * nc):
            y, x = divmod(c, nc)
            win = (slice(y * blk, (y + 1) * blk), slice(x * blk, (x + 1) * blk))
            ca[win] = _embed_one(ca[win], bits[c], scale)
        yuv[:r4, :c4, ch] = idwt2_haar((ca, hvd))
    return yuv


def decode_per_block(yuv, scales=DEFAULT_SCALES, blk=DEFAULT_BLK):
    rows, cols, _ = yuv.shape
    block_num = rows * cols // 4 // (blk * blk)
    wm_bits = np.zeros((3, block_num))
    r4, c4 = rows // 4 * 4, cols // 4 * 4
    for ch in range(3):
        scale = scales[ch]
        if scale <= 0:
            continue
        ca, _ = dwt2_haar(yuv[:r4, :c4, ch])
        nr, nc = ca.shape[0] // blk, ca.shape[1] // blk
        for c in range(nr * nc):
            y, x = divmod(c, nc)
            wm_bits[ch, c] = _extract_one(ca[y * blk:(y + 1) * blk, x * blk:(x + 1) * blk], scale)
    return wm_bits[1].reshape(1, -1)
