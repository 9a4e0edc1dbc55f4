"""B200 drop-in for ``offmark.degenerator.de_shuffler`` (src/offmark/degenerator/de_shuffler.py)."""
import numpy as np
import torch

from b200wm import ops
from .._frames import device_of


class DeShuffler:
    """``set_shape`` / ``degenerate`` as in the reference (de_shuffler.py:8-22).  The per-position
    counting runs in ``b200wm_vote_counts`` and the float64 mean / threshold / un-permute in
    ``b200wm_vote_finish``.  The vote is always taken on the values of the array that is passed in (a slice, a
    copy or an edited array gives the vote of THAT array, like the reference); frame-rate callers use
    ``degenerate_counts`` on the counts the extract kernels leave on the GPU."""

    def __init__(self, key=None, device=None):
        self.key = key
        self.device = device

    def set_shape(self, payload_shape):
        self.payload_len = int(np.array(payload_shape).prod())
        self.payload_idx = np.arange(self.payload_len)
        np.random.RandomState(self.key).shuffle(self.payload_idx)     # host: MT19937 parity
        self._perm = None
        return self

    def _perm_on(self, device):
        if self._perm is None or self._perm.device != device:
            self._perm = torch.from_numpy(self.payload_idx.astype(np.int32)).to(device)
        return self._perm

    def _patterns(self, wm):
        flat = np.asarray(wm).flatten()
        block_num = flat.size
        packed, _ = ops.pack_bits(flat, device=device_of(self.device))
        counts = ops.vote_counts(packed, block_num, self.payload_len)
        patterns, _ = ops.vote_finish(counts, block_num, self._perm_on(packed.device))
        return patterns

    def degenerate(self, wm):
        return self._patterns(wm)[0].cpu().numpy()

    def format_pattern(self, pattern):
        """What ``degenerate`` returns for one voted pattern (uint8 ``[payload_len]``); the batched drivers
        pass every pattern of ``degenerate_counts`` through it."""
        return pattern

    def degenerate_counts(self, pos_counts, block_num):
        """Batched form: int32 ``[N, payload_len]`` counts on the GPU ->
        (patterns uint8 [N, L], packed int64 [N] or None), left on the GPU."""
        return ops.vote_finish(pos_counts, block_num, self._perm_on(pos_counts.device))
