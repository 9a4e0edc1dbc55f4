"""B200 drop-in for ``offmark.generator.shuffler`` (src/offmark/generator/shuffler.py).

Stays on the host: the permutation must come from numpy's MT19937 ``RandomState(key)`` to
match the reference bit for bit (shuffler.py:22), and the result is ``capacity`` bits."""
import math

import numpy as np


class Shuffler:

    def __init__(self, key=None):
        self.key = key

    @staticmethod
    def wm_type():
        return "bits"

    def generate_wm(self, payload, capacity):
        """Payload permuted by ``RandomState(key).shuffle`` and repeated to ``capacity``
        (shuffler.py:15-25): ``wm.flat[c] == shuffled[c % len(payload)]``."""
        total = int(np.prod(capacity))
        shuffled = np.array(payload).copy()
        np.random.RandomState(self.key).shuffle(shuffled)
        repeats = int(math.ceil(total / shuffled.size))
        return np.tile(shuffled.reshape(1, -1), (repeats, 1)).reshape(-1)[:total].reshape(capacity)
