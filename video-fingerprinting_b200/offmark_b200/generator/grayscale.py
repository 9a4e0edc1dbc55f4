"""B200 drop-in for ``offmark.generator.grayscale`` (src/offmark/generator/grayscale.py)."""
import warnings

import numpy as np

from .shuffler import Shuffler


class GrayScale:

    def __init__(self, key=None):
        self.key = key

    @staticmethod
    def wm_type():
        return "grayscale"

    def generate_wm(self, payload, capacity):
        """Image pixels thresholded at 127 (grayscale.py:27), then shuffled and repeated like
        ``Shuffler`` (:28-31)."""
        total = int(np.prod(capacity))
        image = np.asarray(payload)
        if image.size > total:
            warnings.warn("\nImage size {0} is greater than the embed's capacity: {1} pixels".format(image.shape, total),
                          stacklevel=3)
        bits = (image > 127).astype(np.uint8).flatten()
        return Shuffler(key=self.key).generate_wm(bits, capacity)
