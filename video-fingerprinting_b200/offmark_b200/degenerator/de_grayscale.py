"""B200 drop-in for ``offmark.degenerator.de_grayscale`` (src/offmark/degenerator/de_grayscale.py)."""
import numpy as np

from .de_shuffler import DeShuffler


class DeGrayScale(DeShuffler):
    """Same vote as ``DeShuffler`` over ``prod(shape)`` positions, returned as a 0/255 image
    (de_grayscale.py:15-23)."""

    def set_shape(self, payload_shape):
        self.payload_shape = payload_shape
        return super().set_shape(payload_shape)

    def format_pattern(self, pattern):
        return (np.asarray(pattern).astype(np.uint8) * 255).reshape(self.payload_shape)

    def degenerate(self, wm_bits):
        return self.format_pattern(self._patterns(wm_bits)[0].cpu().numpy())
