"""B200 drop-in for ``offmark.extract.dct_decoder`` (src/offmark/extract/dct_decoder.py)."""
import numpy as np

from b200wm import ops
from .._frames import FrameOnDevice
from .._dct_masks import DctMasks


class DctDecoder(DctMasks):
    """Same constructor, ``decode`` and public mask methods as the reference class (dct_decoder.py:4-89)."""

    def __init__(self, key=None, alpha=20, device=None):
        self.key = key
        self.alpha = alpha
        self.device = device

    def decode(self, yuv):
        frame = FrameOnDevice(yuv, self.device)
        rows, cols, _ = frame.dev.shape
        block_num = rows * cols // 8 // 8
        raw, _ = ops.dct8_decode(frame.dev, frame.dev, alpha=self.alpha, lum_channel=0, channel=1)
        return ops.unpack_bits(raw, block_num).astype(np.float64).reshape(1, -1)
