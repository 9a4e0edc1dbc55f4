"""Batched, device-resident entry points over the C ABI (torch tensors in, torch tensors out).

PyTorch is plumbing here: it owns device memory and the current stream, and every
function below hands raw pointers to libb200wm.so.  Nothing in this module computes
on the CPU or with torch operators; without the library or a CUDA device it raises.

Tensor conventions
  planes      uint8 or float32 CUDA tensor, ``[N, H, W]`` / ``[H, W]`` (planar) or
              ``[N, H, W, C]`` / ``[H, W, C]`` with ``channel=c`` (interleaved).  Any strides
              are fine as long as they are positive; the fast path needs planar uint8 with
              8-byte aligned rows.
  packed bits int32 tensors holding little-endian uint32 bit words (bit c -> word c>>5, bit c&31).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, Plane

_ESIZE = {torch.uint8: (1, _lib.U8), torch.float32: (4, _lib.F32)}
_VALIDATE = os.environ.get("B200WM_VALIDATE", "0") not in ("", "0")


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.B200wmError("no CUDA device: the b200wm kernels have no CPU fallback")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    """Device pointer of a tensor (None -> NULL).  Empty tensors made by ``_empty`` still have a
    backing allocation, so the library never sees NULL for a required buffer."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() if t.numel() else t.untyped_storage().data_ptr())


def _empty(shape, dtype, device):
    """torch.empty that keeps a non-NULL pointer even when a dimension is zero."""
    n = 1
    for d in shape:
        n *= d
    if n > 0:
        return torch.empty(shape, dtype=dtype, device=device)
    return torch.empty((1,), dtype=dtype, device=device)[:0].reshape(shape)


def _as_nhw(t, channel):
    """View ``t`` as [N, H, W] (no copy)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    if t.dtype not in _ESIZE:
        raise ValueError(f"unsupported dtype {t.dtype}: uint8 or float32 expected")
    if channel is not None:
        if t.dim() not in (3, 4):
            raise ValueError("interleaved frames must be [H, W, C] or [N, H, W, C]")
        t = t[..., channel]
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3:
        raise ValueError("planes must be [H, W] or [N, H, W] (or interleaved with channel=...)")
    return t


def describe(t, channel=None):
    """(view, b200wm_plane) for a tensor of planes."""
    v = _as_nhw(t, channel)
    esize, code = _ESIZE[v.dtype]
    n, h, w = v.shape
    sn, sh, sw = v.stride()
    if n == 1:
        sn = 0
    if min(sh, sw) <= 0 or sn < 0:
        raise ValueError("planes need positive strides")
    return v, Plane(code, n, h, w, sh * esize, sn * esize, sw, 0)


def geometry(height, width):
    """(block_num, tile_count, words_per_frame) - the reference's capacity rules."""
    return (int(lib.b200wm_block_num(height, width)), int(lib.b200wm_tile_count(height, width)),
            int(lib.b200wm_words_per_frame(height, width)))


def i420_plane(frames, height, width, which="y"):
    """Strided [N, h, w] view (no copy) of one plane of planar I420 / yuv420p frames - the wire format the
    reference hands to its encoder (video/frame_writer.py:34, pix_fmt='yuv420p').  ``frames`` is a
    uint8 tensor ``[N, width*height*3/2]``; ``which`` is 'y' (height x width), 'u' or 'v' (half size).
    Every DWT/SVD entry point takes the view directly, so a mark can be written into (and read from)
    the luma or a chroma plane of 4:2:0 video without a colour-space round trip."""
    if frames.dim() != 2 or frames.dtype != torch.uint8 or frames.shape[1] != width * height * 3 // 2 or width % 2 or height % 2:
        raise ValueError("frames must be uint8 [N, width*height*3/2] with even width and height")
    n, fb = frames.shape
    stride = frames.stride(0)
    if which == "y":
        return frames.as_strided((n, height, width), (stride, width, 1), frames.storage_offset())
    ch, cw = height // 2, width // 2
    off = width * height + (0 if which == "u" else ch * cw)
    if which not in ("u", "v"):
        raise ValueError("which must be 'y', 'u' or 'v'")
    return frames.as_strided((n, ch, cw), (stride, cw, 1), frames.storage_offset() + off)


# ----------------------------------------------------------------------------- bit packing
def pack_bits(bits, device=None):
    """0/1 array ``[rows, n]`` or ``[n]`` -> (int32 tensor [rows, words], n).  Host-side, tiny."""
    a = np.asarray(bits)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError("bits must be 1-D or 2-D")
    if a.size and not np.isin(a, (0, 1)).all():
        raise ValueError("watermark bits must be 0 or 1")
    rows, n = a.shape
    words = max(1, (n + 31) // 32)
    packed = np.zeros((rows, words * 4), dtype=np.uint8)
    if n:
        pb = np.packbits(a.astype(np.uint8), axis=1, bitorder="little")
        packed[:, :pb.shape[1]] = pb
    t = torch.from_numpy(packed.view("<i4").copy())
    return (t.to(device) if device is not None else t), n


def unpack_bits(raw_bits, n):
    """int32 tensor [N, words] -> numpy uint8 [N, n] (host)."""
    a = raw_bits.detach().cpu().numpy().view(np.uint8)
    return np.unpackbits(a, axis=1, bitorder="little")[:, :n]


def _check_wm(wm_packed, frame_wm_row, n_frames, validate_rows):
    """The packed watermark table and the per-frame row table must be what the kernels dereference: contiguous
    CUDA int32.  Row values are clamped on the device (never an out-of-bounds read); ``validate_rows`` (or
    B200WM_VALIDATE=1) additionally checks them here, at the price of a device synchronisation."""
    if not isinstance(wm_packed, torch.Tensor) or wm_packed.dtype != torch.int32 or wm_packed.dim() != 2 \
            or not wm_packed.is_contiguous() or not wm_packed.is_cuda:
        raise ValueError("wm_packed must be a contiguous CUDA int32 [rows, words] tensor (see pack_bits)")
    if frame_wm_row is not None:
        if not isinstance(frame_wm_row, torch.Tensor) or frame_wm_row.dtype != torch.int32 or frame_wm_row.numel() != n_frames \
                or not frame_wm_row.is_cuda or not frame_wm_row.is_contiguous():
            raise ValueError("frame_wm_row must be a contiguous CUDA int32 [n_frames] tensor")
        if (validate_rows or _VALIDATE) and frame_wm_row.numel():
            lo, hi = int(frame_wm_row.min()), int(frame_wm_row.max())
            if lo < 0 or hi >= wm_packed.shape[0]:
                raise ValueError(f"frame_wm_row holds rows {lo}..{hi} but wm_packed has {wm_packed.shape[0]} rows")


# ----------------------------------------------------------------------------- DWT / SVD pair
def dwtsvd_embed_(planes, wm_packed, wm_len, scale=15.0, channel=None, frame_wm_row=None, out=None, validate_rows=False):
    """In-place (or into ``out``, same geometry, pre-filled) embed.  Returns the written tensor."""
    require_cuda()
    v, pl = describe(planes, channel)
    dst_t = planes if out is None else out
    dv, dpl = describe(dst_t, channel)
    if (dpl.pitch_bytes, dpl.frame_stride_bytes, dpl.elem_stride, dpl.height, dpl.width, dpl.n_frames, dpl.dtype) != \
            (pl.pitch_bytes, pl.frame_stride_bytes, pl.elem_stride, pl.height, pl.width, pl.n_frames, pl.dtype):
        raise ValueError("out must have the geometry of planes")
    _check_wm(wm_packed, frame_wm_row, pl.n_frames, validate_rows)
    check(lib.b200wm_dwtsvd_embed(_ptr(v), _ptr(dv), C.byref(pl), _ptr(wm_packed), wm_packed.shape[0], wm_packed.shape[1],
                                  int(wm_len), _ptr(frame_wm_row), float(scale), _stream()))
    return dst_t


def dwtsvd_embed_copies(planes, wm_packed, wm_len, n_copies, scale=15.0, copy_wm_row=None, out=None):
    """One read, ``n_copies`` marked copies of every uint8 plane (tests/mark_video_to_hls.py:330-354).

    ``out`` (optional) is a uint8 tensor ``[n_copies, N, ...]`` whose slices ``out[c]`` have the strides of
    ``planes``; by default it is allocated and pre-filled with the source so that samples outside the
    walked tiles (edges not covered by 8x8 tiles, chroma planes of an I420 buffer) are copies too.
    Copy ``c`` of frame ``f`` carries watermark row ``copy_wm_row[f, c]`` (default: row ``c``)."""
    require_cuda()
    v, pl = describe(planes)
    if v.dtype != torch.uint8:
        raise ValueError("multi-copy embed takes uint8 planes")
    if wm_packed.dtype != torch.int32 or wm_packed.dim() != 2 or not wm_packed.is_contiguous() or not wm_packed.is_cuda:
        raise ValueError("wm_packed must be a contiguous CUDA int32 [rows, words] tensor (see pack_bits)")
    n_copies = int(n_copies)
    if out is None:
        # same strides as the source inside each copy (the library addresses copy c with the source's pitches)
        extent = sum((d - 1) * st for d, st in zip(planes.shape, planes.stride())) + 1
        out = torch.empty_strided((n_copies,) + tuple(planes.shape), ((extent + 15) // 16 * 16,) + tuple(planes.stride()),
                                  dtype=torch.uint8, device=planes.device)
        out.copy_(planes.unsqueeze(0).expand_as(out))
    if out.dtype != torch.uint8 or not out.is_cuda or out.shape[0] != n_copies or out.dim() != planes.dim() + 1:
        raise ValueError("out must be a CUDA uint8 tensor [n_copies, *planes.shape]")
    if n_copies:
        ov, opl = describe(out[0])
        if (opl.pitch_bytes, opl.frame_stride_bytes, opl.elem_stride, opl.height, opl.width, opl.n_frames) != \
                (pl.pitch_bytes, pl.frame_stride_bytes, pl.elem_stride, pl.height, pl.width, pl.n_frames):
            raise ValueError("every out[c] must have the geometry (strides included) of planes")
    if copy_wm_row is not None and (copy_wm_row.dtype != torch.int32 or tuple(copy_wm_row.shape) != (pl.n_frames, n_copies)
                                    or not copy_wm_row.is_cuda or not copy_wm_row.is_contiguous()):
        raise ValueError("copy_wm_row must be a contiguous CUDA int32 [n_frames, n_copies] tensor")
    if n_copies == 0:
        return out
    check(lib.b200wm_dwtsvd_embed_copies(_ptr(v), C.byref(pl), _ptr(out), int(out.stride(0)) if n_copies else 0, n_copies,
                                         _ptr(wm_packed), wm_packed.shape[0], wm_packed.shape[1], int(wm_len),
                                         _ptr(copy_wm_row), float(scale), _stream()))
    return out


def dwtsvd_extract(planes, scale=15.0, payload_len=None, channel=None, raw_bits=None, pos_counts=None):
    """-> (raw_bits int32 [N, words], pos_counts int32 [N, payload_len] or None)."""
    require_cuda()
    v, pl = describe(planes, channel)
    _, _, words = geometry(pl.height, pl.width)
    if raw_bits is None:
        raw_bits = _empty((pl.n_frames, words), torch.int32, v.device)
    if payload_len is not None and pos_counts is None:
        pos_counts = torch.empty((pl.n_frames, payload_len), dtype=torch.int32, device=v.device)
    check(lib.b200wm_dwtsvd_extract(_ptr(v), C.byref(pl), float(scale), _ptr(raw_bits),
                                    words, int(payload_len or 0), _ptr(pos_counts), _stream()))
    return raw_bits, pos_counts


def dwtsvd_sigma(planes, channel=None):
    """sigma_0 of every walked block, float32 [N, tiles] (validation aid)."""
    require_cuda()
    v, pl = describe(planes, channel)
    _, tiles, _ = geometry(pl.height, pl.width)
    sigma = _empty((pl.n_frames, tiles), torch.float32, v.device)
    check(lib.b200wm_dwtsvd_sigma(_ptr(v), C.byref(pl), _ptr(sigma), _stream()))
    return sigma


def dwtsvd_sigma_dct(planes, channel=None):
    """sigma_0 computed WITH the 4x4 block DCT, as the reference writes it (validation aid)."""
    require_cuda()
    v, pl = describe(planes, channel)
    _, tiles, _ = geometry(pl.height, pl.width)
    sigma = _empty((pl.n_frames, tiles), torch.float32, v.device)
    check(lib.b200wm_dwtsvd_sigma_dct(_ptr(v), C.byref(pl), _ptr(sigma), _stream()))
    return sigma


# ----------------------------------------------------------------------------- 8x8 DCT pair
def dct8_masks(lum, channel=None):
    """-> (block_mean f32 [N, nb], tex_mask f32 [N, nb], frame_sum f64 [N])."""
    require_cuda()
    v, pl = describe(lum, channel)
    nb = (pl.height // 8) * (pl.width // 8)
    block_mean = _empty((pl.n_frames, nb), torch.float32, v.device)
    tex = _empty((pl.n_frames, nb), torch.float32, v.device)
    frame_sum = _empty((pl.n_frames,), torch.float64, v.device)
    check(lib.b200wm_dct8_masks(_ptr(v), C.byref(pl), _ptr(block_mean), _ptr(tex), _ptr(frame_sum), _stream()))
    return block_mean, tex, frame_sum


def dct8_embed_(planes, masks, wm_packed, wm_len, alpha=20.0, channel=None, frame_wm_row=None, validate_rows=False):
    require_cuda()
    v, pl = describe(planes, channel)
    block_mean, tex, frame_sum = masks
    _check_wm(wm_packed, frame_wm_row, pl.n_frames, validate_rows)
    check(lib.b200wm_dct8_embed(_ptr(v), _ptr(v), C.byref(pl), _ptr(block_mean), _ptr(tex), _ptr(frame_sum),
                                _ptr(wm_packed), wm_packed.shape[0], wm_packed.shape[1], int(wm_len), _ptr(frame_wm_row),
                                float(alpha), _stream()))
    return planes


def dct8_extract(planes, masks, alpha=20.0, payload_len=None, channel=None):
    require_cuda()
    v, pl = describe(planes, channel)
    block_mean, tex, frame_sum = masks
    _, _, words = geometry(pl.height, pl.width)
    raw_bits = _empty((pl.n_frames, words), torch.int32, v.device)
    pos_counts = None
    if payload_len is not None:
        pos_counts = torch.empty((pl.n_frames, payload_len), dtype=torch.int32, device=v.device)
    check(lib.b200wm_dct8_extract(_ptr(v), C.byref(pl), _ptr(block_mean), _ptr(tex), _ptr(frame_sum), float(alpha),
                                  _ptr(raw_bits), words, int(payload_len or 0), _ptr(pos_counts), _stream()))
    return raw_bits, pos_counts


def dct8_encode_(lum, planes, wm_packed, wm_len, alpha=20.0, lum_channel=None, channel=None, frame_wm_row=None, validate_rows=False):
    """``DctEncoder.encode`` in one call, no mask arrays: masks from ``lum`` (channel ``lum_channel`` of interleaved
    frames) and the quantiser on ``planes`` (channel ``channel``), in place.  Returns ``planes``."""
    require_cuda()
    lv, lpl = describe(lum, lum_channel)
    v, pl = describe(planes, channel)
    _check_wm(wm_packed, frame_wm_row, pl.n_frames, validate_rows)
    frame_sum = _empty((pl.n_frames,), torch.float64, v.device)
    check(lib.b200wm_dct8_encode(_ptr(lv), C.byref(lpl), _ptr(v), _ptr(v), C.byref(pl), _ptr(frame_sum), _ptr(wm_packed),
                                 wm_packed.shape[0], wm_packed.shape[1], int(wm_len), _ptr(frame_wm_row), float(alpha), _stream()))
    return planes


def dct8_decode(lum, planes, alpha=20.0, payload_len=None, lum_channel=None, channel=None):
    """``DctDecoder.decode`` in one call -> (raw_bits int32 [N, words], pos_counts int32 [N, payload_len] or None)."""
    require_cuda()
    lv, lpl = describe(lum, lum_channel)
    v, pl = describe(planes, channel)
    _, _, words = geometry(pl.height, pl.width)
    raw_bits = _empty((pl.n_frames, words), torch.int32, v.device)
    pos_counts = torch.empty((pl.n_frames, payload_len), dtype=torch.int32, device=v.device) if payload_len is not None else None
    frame_sum = _empty((pl.n_frames,), torch.float64, v.device)
    check(lib.b200wm_dct8_decode(_ptr(lv), C.byref(lpl), _ptr(v), C.byref(pl), _ptr(frame_sum), float(alpha), _ptr(raw_bits), words,
                                 int(payload_len or 0), _ptr(pos_counts), _stream()))
    return raw_bits, pos_counts


# ----------------------------------------------------------------------------- votes
def vote_counts(raw_bits, block_num, payload_len):
    require_cuda()
    if raw_bits.dtype != torch.int32 or raw_bits.dim() != 2 or not raw_bits.is_cuda or not raw_bits.is_contiguous():
        raise ValueError("raw_bits must be a contiguous CUDA int32 [N, words] tensor")
    n, words = raw_bits.shape
    if int(payload_len) <= 0 or not 0 <= int(block_num) <= 32 * words:
        raise ValueError("payload_len must be positive and block_num within the packed words")
    counts = torch.empty((n, payload_len), dtype=torch.int32, device=raw_bits.device)
    check(lib.b200wm_vote_counts(_ptr(raw_bits), n, words, int(block_num), int(payload_len), _ptr(counts), _stream()))
    return counts


def vote_finish(pos_counts, block_num, perm):
    """Per-frame finish of DeShuffler.degenerate -> (patterns uint8 [N, L], packed int64 [N] or None)."""
    require_cuda()
    if pos_counts.dtype != torch.int32 or pos_counts.dim() != 2 or not pos_counts.is_cuda or not pos_counts.is_contiguous():
        raise ValueError("pos_counts must be a contiguous CUDA int32 [N, payload_len] tensor")
    n, length = pos_counts.shape
    if perm.dtype != torch.int32 or perm.numel() != length or not perm.is_cuda or not perm.is_contiguous():
        raise ValueError("perm must be a contiguous CUDA int32 tensor of payload_len entries")
    patterns = torch.empty((n, length), dtype=torch.uint8, device=pos_counts.device)
    packed = torch.empty((n,), dtype=torch.int64, device=pos_counts.device) if length <= 64 else None
    check(lib.b200wm_vote_finish(_ptr(pos_counts), n, length, int(block_num), _ptr(perm), _ptr(patterns), _ptr(packed),
                                 _stream()))
    return patterns, packed


INT32_MAX = 2 ** 31 - 1


def _check_hist_inputs(packed, payload_len, frame_segment, frame_order):
    """The histogram kernels read 8 bytes per frame from ``packed`` and 4 from the per-frame tables."""
    if not isinstance(packed, torch.Tensor) or not packed.is_cuda or packed.dtype != torch.int64 or packed.dim() != 1 \
            or not packed.is_contiguous():
        raise ValueError("packed must be a contiguous 1-D CUDA int64 tensor (the second result of vote_finish)")
    if not 1 <= int(payload_len) <= 16:
        raise ValueError("payload_len must be 1..16 for the pattern histogram (gathered_pattern_vote takes any length)")
    for name, t in (("frame_segment", frame_segment), ("frame_order", frame_order)):
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or t.device != packed.device or t.dtype != torch.int32 or not t.is_contiguous() \
                or t.numel() != packed.numel():
            raise ValueError(f"{name} must be a contiguous int32 tensor on packed's device with one entry per frame")


def pattern_hist(packed, payload_len, n_segments=1, frame_segment=None, frame_order=None, order_offset=0, state=None):
    """Accumulate the device half of the cross-frame vote.  ``state`` (from a previous call) is
    updated in place.  -> dict(hist, first_seen, bit_votes, seg_frames).  Frames whose segment is outside
    ``[0, n_segments)`` or whose pattern has bits above ``payload_len`` are ignored."""
    require_cuda()
    _check_hist_inputs(packed, payload_len, frame_segment, frame_order)
    dev = packed.device
    if state is None:
        state = {
            "hist": torch.zeros((n_segments, 1 << payload_len), dtype=torch.int32, device=dev),
            "first_seen": torch.full((n_segments, 1 << payload_len), INT32_MAX, dtype=torch.int32, device=dev),
            "bit_votes": torch.zeros((n_segments, payload_len), dtype=torch.int32, device=dev),
            "seg_frames": torch.zeros((n_segments,), dtype=torch.int32, device=dev),
        }
    check(lib.b200wm_pattern_hist(_ptr(packed), _ptr(frame_segment), _ptr(frame_order), int(order_offset),
                                  packed.numel(), int(payload_len), int(n_segments), _ptr(state["hist"]),
                                  _ptr(state["first_seen"]), _ptr(state["bit_votes"]), _ptr(state["seg_frames"]),
                                  _stream()))
    return state


def pattern_hist_publish(packed, payload_len, n_segments, state, frame_segment, frame_order, order_offset, peers_dev, block_len,
                         world, rank, epoch, ticket, status):
    """``pattern_hist`` into ``state`` (views of this rank's block) fused with the NVLink exchange of the block
    (``b200wm_pattern_hist_publish``).  ``peers_dev``: device address of the array of peer base pointers."""
    require_cuda()
    _check_hist_inputs(packed, payload_len, frame_segment, frame_order)
    check(lib.b200wm_pattern_hist_publish(_ptr(packed), _ptr(frame_segment), _ptr(frame_order), int(order_offset), packed.numel(),
                                          int(payload_len), int(n_segments), _ptr(state["hist"]), _ptr(state["first_seen"]),
                                          _ptr(state["bit_votes"]), _ptr(state["seg_frames"]), C.c_void_p(int(peers_dev)),
                                          int(block_len), int(world), int(rank), int(epoch) & 0xFFFFFFFF, _ptr(ticket), _ptr(status),
                                          _stream()))
    return state


def vote_exchange_wait(peers_dev, block_len, world, rank, epoch, status):
    """Enqueue the wait for every peer's block of exchange ``epoch`` (``b200wm_vote_exchange_wait``)."""
    require_cuda()
    check(lib.b200wm_vote_exchange_wait(C.c_void_p(int(peers_dev)), int(block_len), int(world), int(rank), int(epoch) & 0xFFFFFFFF,
                                        _ptr(status), _stream()))


def vote_state_reset(flat, n_zero):
    """``flat`` int32 CUDA tensor: entries [0, n_zero) <- 0, the rest <- INT32_MAX, one launch."""
    require_cuda()
    if flat.dtype != torch.int32 or not flat.is_cuda or not flat.is_contiguous():
        raise ValueError("state must be a contiguous CUDA int32 tensor")
    check(lib.b200wm_vote_state_reset(_ptr(flat), int(n_zero), flat.numel(), _stream()))
    return flat


# ----------------------------------------------------------------------------- colour bracket
def bgr8_to_yuv32(frames):
    """uint8 [..., 3] contiguous -> float32 same shape (video/embedder.py:34)."""
    require_cuda()
    if frames.dtype != torch.uint8 or not frames.is_contiguous() or frames.shape[-1] != 3:
        raise ValueError("frames must be contiguous uint8 [..., 3]")
    out = torch.empty(frames.shape, dtype=torch.float32, device=frames.device)
    check(lib.b200wm_bgr8_to_yuv32(_ptr(frames), _ptr(out), frames.numel() // 3, _stream()))
    return out


def yuv32_to_bgr8(yuv):
    """float32 [..., 3] contiguous -> uint8 with clip / round-half-even (video/embedder.py:36-38)."""
    require_cuda()
    if yuv.dtype != torch.float32 or not yuv.is_contiguous() or yuv.shape[-1] != 3:
        raise ValueError("yuv must be contiguous float32 [..., 3]")
    out = torch.empty(yuv.shape, dtype=torch.uint8, device=yuv.device)
    check(lib.b200wm_yuv32_to_bgr8(_ptr(yuv), _ptr(out), yuv.numel() // 3, _stream()))
    return out


# ----------------------------------------------------------------------------- fused colour bracket
def _rgb_frames(frames):
    if not isinstance(frames, torch.Tensor) or not frames.is_cuda or frames.dtype != torch.uint8:
        raise ValueError("frames must be a CUDA uint8 tensor")
    f = frames.unsqueeze(0) if frames.dim() == 3 else frames
    if f.dim() != 4 or f.shape[3] != 3 or f.stride(3) != 1 or f.stride(2) != 3:
        raise ValueError("frames must be [N, H, W, 3] (or [H, W, 3]) with packed pixels")
    n, h, w, _ = f.shape
    return f, n, h, w, f.stride(1), (f.stride(0) if n > 1 else 0)


def dwtsvd_embed_rgb8_(frames, wm_packed, wm_len, scales=(0.0, 15.0, 0.0), frame_wm_row=None, validate_rows=False):
    """uint8 [N, H, W, 3] frames marked in place: colour conversion, DWT/SVD embed on the channels
    with a positive scale and conversion back, in one kernel (video/embedder.py:33-39)."""
    require_cuda()
    f, n, h, w, pitch, fstride = _rgb_frames(frames)
    sc = (C.c_float * 3)(*[float(s) for s in scales])
    _check_wm(wm_packed, frame_wm_row, n, validate_rows)
    check(lib.b200wm_dwtsvd_embed_rgb8(_ptr(f), _ptr(f), n, h, w, pitch, fstride, sc, _ptr(wm_packed), wm_packed.shape[0],
                                       wm_packed.shape[1], int(wm_len), _ptr(frame_wm_row), _stream()))
    return frames


def dwtsvd_extract_rgb8(frames, scale=15.0, channel=1, payload_len=None):
    """uint8 [N, H, W, 3] frames -> (raw_bits, pos_counts) of YUV channel ``channel`` (video/extractor.py:30-33)."""
    require_cuda()
    f, n, h, w, pitch, fstride = _rgb_frames(frames)
    _, _, words = geometry(h, w)
    raw_bits = _empty((n, words), torch.int32, f.device)
    pos_counts = torch.empty((n, payload_len), dtype=torch.int32, device=f.device) if payload_len is not None else None
    check(lib.b200wm_dwtsvd_extract_rgb8(_ptr(f), n, h, w, pitch, fstride, int(channel), float(scale), _ptr(raw_bits), words,
                                         int(payload_len or 0), _ptr(pos_counts), _stream()))
    return raw_bits, pos_counts


# ----------------------------------------------------------------------------- host-buffer entry points
def _host_planes(t):
    """CPU uint8 tensor [N, H, W] (any positive strides with contiguous rows) -> (tensor, plane)."""
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.uint8 or t.dim() != 3 or t.stride(2) != 1:
        raise ValueError("host planes must be a CPU uint8 [N, H, W] array with contiguous rows")
    n, h, w = t.shape
    return t, Plane(_lib.U8, n, h, w, t.stride(1), t.stride(0) if n > 1 else 0, 1, 0)


def dwtsvd_mark_host(src, dst, wm_rows, scale=15.0, frame_wm_row=None, chunk_frames=0, wm_len=None):
    """Mark host-resident planes: ``src``/``dst`` CPU uint8 ``[N, H, W]`` (pinned for full PCIe speed; may be
    the same array), ``wm_rows`` 0/1 array ``[rows, n]`` - or, with ``wm_len`` given, rows already packed by
    ``pack_bits`` (CPU int32 ``[rows, words]``).  One C-ABI call; copies and kernels overlap inside."""
    require_cuda()
    s, pl = _host_planes(src)
    d, dpl = _host_planes(dst)
    if (dpl.height, dpl.width, dpl.n_frames, dpl.pitch_bytes, dpl.frame_stride_bytes) != \
            (pl.height, pl.width, pl.n_frames, pl.pitch_bytes, pl.frame_stride_bytes):
        raise ValueError("dst must have the geometry of src")
    if wm_len is None:
        packed, n = pack_bits(wm_rows)
    else:
        packed, n = wm_rows, int(wm_len)
        if not isinstance(packed, torch.Tensor) or packed.is_cuda or packed.dtype != torch.int32 or packed.dim() != 2 \
                or not packed.is_contiguous():
            raise ValueError("packed watermark rows must be a contiguous CPU int32 [rows, words] tensor")
    rows = None
    if frame_wm_row is not None:
        rows = torch.as_tensor(frame_wm_row, dtype=torch.int32).contiguous()
        if rows.numel() != pl.n_frames:
            raise ValueError("frame_wm_row needs one entry per frame")
    check(lib.b200wm_dwtsvd_mark_host(C.c_void_p(s.data_ptr()), C.c_void_p(d.data_ptr()), C.byref(pl), C.c_void_p(packed.data_ptr()),
                                      packed.shape[0], packed.shape[1], int(n),
                                      C.c_void_p(rows.data_ptr()) if rows is not None else None, float(scale), int(chunk_frames)))
    return dst


def dwtsvd_detect_host(src, perm, scale=15.0, chunk_frames=0, want_raw_bits=False):
    """Extract + per-frame vote of host-resident planes -> patterns uint8 ``[N, L]`` (numpy)
    [, raw bits uint32 ``[N, words]``]."""
    require_cuda()
    s, pl = _host_planes(src)
    perm = torch.as_tensor(np.asarray(perm), dtype=torch.int32).contiguous()
    length = perm.numel()
    _, _, words = geometry(pl.height, pl.width)
    patterns = torch.empty((pl.n_frames, length), dtype=torch.uint8)
    raw = torch.empty((pl.n_frames, max(words, 1)), dtype=torch.int32) if want_raw_bits else None
    check(lib.b200wm_dwtsvd_detect_host(C.c_void_p(s.data_ptr()), C.byref(pl), float(scale), length, C.c_void_p(perm.data_ptr()),
                                        C.c_void_p(patterns.data_ptr()), C.c_void_p(raw.data_ptr()) if raw is not None else None,
                                        None, int(chunk_frames)))
    if want_raw_bits:
        return patterns.numpy(), raw[:, :words].numpy().view(np.uint32)
    return patterns.numpy()


def dwtsvd_mark_verify_host(src, dst, wm_rows, perm, scale=15.0, frame_wm_row=None, chunk_frames=0, wm_len=None,
                            want_raw_bits=False):
    """Mark host-resident planes AND read the payload back from the marked planes in the same pass (one upload, one
    download per frame): the reference's mark-then-verify step (tests/mark_video_to_hls.py:356-399).  Arguments as
    ``dwtsvd_mark_host`` plus the de-shuffling permutation; -> patterns uint8 ``[N, L]`` (numpy) [, raw bits]."""
    require_cuda()
    s, pl = _host_planes(src)
    d, dpl = _host_planes(dst)
    if (dpl.height, dpl.width, dpl.n_frames, dpl.pitch_bytes, dpl.frame_stride_bytes) != \
            (pl.height, pl.width, pl.n_frames, pl.pitch_bytes, pl.frame_stride_bytes):
        raise ValueError("dst must have the geometry of src")
    if wm_len is None:
        packed, n = pack_bits(wm_rows)
    else:
        packed, n = wm_rows, int(wm_len)
        if not isinstance(packed, torch.Tensor) or packed.is_cuda or packed.dtype != torch.int32 or packed.dim() != 2 \
                or not packed.is_contiguous():
            raise ValueError("packed watermark rows must be a contiguous CPU int32 [rows, words] tensor")
    rows = None
    if frame_wm_row is not None:
        rows = torch.as_tensor(frame_wm_row, dtype=torch.int32).contiguous()
        if rows.numel() != pl.n_frames:
            raise ValueError("frame_wm_row needs one entry per frame")
    perm = torch.as_tensor(np.asarray(perm), dtype=torch.int32).contiguous()
    length = perm.numel()
    _, _, words = geometry(pl.height, pl.width)
    patterns = torch.empty((pl.n_frames, length), dtype=torch.uint8)
    raw = torch.empty((pl.n_frames, max(words, 1)), dtype=torch.int32) if want_raw_bits else None
    check(lib.b200wm_dwtsvd_mark_verify_host(C.c_void_p(s.data_ptr()), C.c_void_p(d.data_ptr()), C.byref(pl),
                                             C.c_void_p(packed.data_ptr()), packed.shape[0], packed.shape[1], int(n),
                                             C.c_void_p(rows.data_ptr()) if rows is not None else None, float(scale), length,
                                             C.c_void_p(perm.data_ptr()), C.c_void_p(patterns.data_ptr()),
                                             C.c_void_p(raw.data_ptr()) if raw is not None else None, None, int(chunk_frames)))
    if want_raw_bits:
        return patterns.numpy(), raw[:, :words].numpy().view(np.uint32)
    return patterns.numpy()


def host_scratch_release():
    """Free the streams and device scratch that the host-buffer entry points keep between calls."""
    check(lib.b200wm_host_scratch_release())


# ----------------------------------------------------------------------------- distortion channel
def attack_jpeg_requant_(planes, quality):
    """In-place JPEG-like requantisation of planar uint8 planes (BASELINE config 5)."""
    require_cuda()
    v, pl = describe(planes)
    check(lib.b200wm_attack_jpeg_requant(_ptr(v), _ptr(v), C.byref(pl), int(quality), _stream()))
    return planes


def attack_add_noise_(planes, noise):
    """In-place ``clip(round(x + noise))``; ``noise`` float32 ``[N, H, W]`` contiguous on the same device."""
    require_cuda()
    v, pl = describe(planes)
    if noise.dtype != torch.float32 or not noise.is_contiguous() or noise.numel() != v.numel():
        raise ValueError("noise must be a contiguous float32 tensor with one value per sample")
    check(lib.b200wm_attack_add_noise(_ptr(v), _ptr(v), C.byref(pl), _ptr(noise), _stream()))
    return planes


INTER_LINEAR, INTER_AREA = 1, 3      # cv2's values


def attack_resize(planes, size, interpolation):
    """cv2.resize(plane, (width, height), interpolation=...) for every uint8 plane of ``[N, H, W]``;
    ``size`` is (width, height) like cv2's dsize.  Returns a new tensor."""
    require_cuda()
    v, pl = describe(planes)
    dw, dh = int(size[0]), int(size[1])
    out = _empty((v.shape[0], dh, dw), torch.uint8, v.device)
    _, dpl = describe(out)
    check(lib.b200wm_attack_resize(_ptr(v), C.byref(pl), _ptr(out), C.byref(dpl), int(interpolation), _stream()))
    return out if planes.dim() == 3 else out[0]


def attack_resize_roundtrip_(planes, scale=2.0 / 3.0):
    """In place: INTER_AREA down by ``scale`` then INTER_LINEAR back (oracle/attacks.py:resize_roundtrip)."""
    v, _ = describe(planes)
    h, w = v.shape[1], v.shape[2]
    small = attack_resize(v, (int(round(w * scale)), int(round(h * scale))), INTER_AREA)
    _, pl = describe(small)
    _, dpl = describe(v)
    check(lib.b200wm_attack_resize(_ptr(small), C.byref(pl), _ptr(v), C.byref(dpl), INTER_LINEAR, _stream()))
    return planes


def set_path(path):
    """0 = automatic (TMA-staged kernels when the planes qualify), 1 = vectorised-load kernels only."""
    check(lib.b200wm_set_path(int(path)))


def get_path():
    return int(lib.b200wm_get_path())


def kernel_launches():
    return int(lib.b200wm_kernel_launches())
