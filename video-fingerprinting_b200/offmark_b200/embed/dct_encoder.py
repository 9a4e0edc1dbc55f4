"""B200 drop-in for ``offmark.embed.dct_encoder`` (src/offmark/embed/dct_encoder.py)."""
from b200wm import ops
from .._frames import FrameOnDevice


class DctEncoder:
    """Same constructor, ``read_wm`` / ``wm_capacity`` / ``encode`` as the reference class
    (dct_encoder.py:4-39).  Masks (:41-102) run in ``b200wm_dct8_masks`` on channel 0, the
    block loop (:24-38) in ``b200wm_dct8_embed`` on channel 1."""

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
        masks = ops.dct8_masks(frame.dev, channel=0)
        ops.dct8_embed_(frame.dev, masks, packed, n, alpha=self.alpha, channel=1)
        return frame.write_back()
