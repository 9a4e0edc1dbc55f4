"""B200 drop-in for ``offmark.embed.dwt_dct_svd_encoder`` (src/offmark/embed/dwt_dct_svd_encoder.py)."""
import torch

from b200wm import ops
from .._frames import FrameOnDevice, device_of


class DwtDctSvdEncoder:
    """Same constructor, ``read_wm`` / ``wm_capacity`` / ``encode`` as the reference class
    (dwt_dct_svd_encoder.py:5-27); the per-block loop (:29-45) runs in
    ``b200wm_dwtsvd_embed``.  ``blk`` other than 4 is rejected: the reference's own capacity
    formula (:16) only holds for 4, and the kernel is specialised for it."""

    def __init__(self, key=None, scales=[0, 15, 0], blk=4, device=None):
        if blk != 4:
            raise ValueError("DwtDctSvdEncoder on B200 implements blk=4 only")
        self.key = key
        self.scales = scales
        self.blk = blk
        self.device = device

    def read_wm(self, wm):
        self.wm = wm[0]

    def wm_capacity(self, frame_shape):
        row, col, channels = frame_shape
        return (1, row * col // 64)

    def _packed_wm(self, device):
        return ops.pack_bits(self.wm, device=device)

    def encode(self, yuv):
        """Marks ``yuv`` (float32 H x W x 3, numpy or CUDA tensor) in place and returns it."""
        frame = FrameOnDevice(yuv, self.device)
        packed, n = self._packed_wm(frame.dev.device)
        for channel in range(3):
            if self.scales[channel] <= 0:
                continue
            ops.dwtsvd_embed_(frame.dev, packed, n, scale=self.scales[channel], channel=channel)
        return frame.write_back()

    def mark_rgb8(self, frames):
        """uint8 ``[N, H, W, 3]`` / ``[H, W, 3]`` CUDA frames marked in place in ONE kernel: the colour
        bracket of ``Embedder.__mark_frame`` (video/embedder.py:33-39) fused around ``encode``."""
        packed, n = self._packed_wm(frames.device)
        return ops.dwtsvd_embed_rgb8_(frames, packed, n, scales=self.scales)

    def encode_planes(self, planes, scale=None, frame_wm_row=None, wm_rows=None, out=None):
        """Batched form for device-resident planes (``[N, H, W]`` uint8 or float32): one launch
        for the whole batch.  ``wm_rows`` (2-D 0/1 array) with ``frame_wm_row`` (int32 ``[N]``)
        gives every frame its own watermark row (one row per HLS segment)."""
        dev = device_of(self.device) if not isinstance(planes, torch.Tensor) else planes.device
        packed, n = ops.pack_bits(self.wm if wm_rows is None else wm_rows, device=dev)
        if scale is None:
            scale = max(self.scales)
        return ops.dwtsvd_embed_(planes, packed, n, scale=scale, frame_wm_row=frame_wm_row, out=out)
