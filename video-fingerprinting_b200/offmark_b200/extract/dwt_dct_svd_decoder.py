"""B200 drop-in for ``offmark.extract.dwt_dct_svd_decoder`` (src/offmark/extract/dwt_dct_svd_decoder.py)."""
import numpy as np

from b200wm import ops
from .._frames import FrameOnDevice


class DwtDctSvdDecoder:
    """Same constructor and ``decode`` as the reference class (dwt_dct_svd_decoder.py:5-21); the
    per-block loop (:23-37) runs in ``b200wm_dwtsvd_extract``."""

    def __init__(self, key=None, scales=[0, 15, 0], blk=4, device=None):
        if blk != 4:
            raise ValueError("DwtDctSvdDecoder on B200 implements blk=4 only")
        self.key = key
        self.scales = scales
        self.blk = blk
        self.device = device

    def decode(self, yuv):
        """float32 H x W x 3 (numpy or CUDA tensor) -> float64 (1, rows*cols//64) of 0.0/1.0.
        Like the reference (:21) this is always the row of channel 1: zeros when
        ``scales[1] <= 0``; other channels' bits are computed by the reference and dropped,
        so they are not computed here."""
        frame = FrameOnDevice(yuv, self.device)
        rows, cols, _ = frame.dev.shape
        self.block_num = rows * cols // 4 // (self.blk * self.blk)
        if self.scales[1] <= 0:
            return np.zeros((1, self.block_num))
        raw, _ = ops.dwtsvd_extract(frame.dev, scale=self.scales[1], channel=1)
        return ops.unpack_bits(raw, self.block_num).astype(np.float64).reshape(1, -1)

    def decode_rgb8(self, frame):
        """uint8 ``[H, W, 3]`` CUDA frame -> the same ``(1, N)`` float64 array as ``decode`` of its YUV
        conversion (video/extractor.py:31-32), colour conversion fused into the extract kernel."""
        rows, cols, _ = frame.shape
        self.block_num = rows * cols // 4 // (self.blk * self.blk)
        if self.scales[1] <= 0:
            return np.zeros((1, self.block_num))
        raw, _ = ops.dwtsvd_extract_rgb8(frame, scale=self.scales[1], channel=1)
        return ops.unpack_bits(raw, self.block_num).astype(np.float64).reshape(1, -1)

    def decode_planes(self, planes, scale=None, payload_len=None):
        """Batched form for device-resident planes: -> (raw_bits int32 [N, words],
        pos_counts int32 [N, payload_len] or None), everything left on the GPU."""
        if scale is None:
            scale = max(self.scales)
        return ops.dwtsvd_extract(planes, scale=scale, payload_len=payload_len)
