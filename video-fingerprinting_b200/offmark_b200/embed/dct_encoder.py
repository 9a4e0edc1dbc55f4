"""B200 drop-in for ``offmark.embed.dct_encoder`` (src/offmark/embed/dct_encoder.py)."""
from b200wm import ops
from .._frames import FrameOnDevice
from .._dct_masks import DctMasks


class DctEncoder(DctMasks):
    """Same constructor, ``read_wm`` / ``wm_capacity`` / ``encode`` as the reference class
    (dct_encoder.py:4-39).  ``b200wm_dct8_encode``: the masks of channel 0 (:41-102) and the block
    loop on channel 1 (:24-38) in one call, no mask arrays; ``luminance_mask`` / ``texture_mask`` (:41-102) are there
    for callers that want the masks themselves."""

    def __init__(self, key=None, alpha=20, device=None):
        self.key = key
        self.alpha = alpha
        self.device = device

    def read_wm(self, wm):
        self.wm = wm[0]

    def wm_capacity(self, frame_shape):
        row, col, channels = frame_shape
        return (1, row * col // 64)

    def encode(self, yuv):
        frame = FrameOnDevice(yuv, self.device)
        packed, n = ops.pack_bits(self.wm, device=frame.dev.device)
        ops.dct8_encode_(frame.dev, frame.dev, packed, n, alpha=self.alpha, lum_channel=0, channel=1)
        return frame.write_back()
