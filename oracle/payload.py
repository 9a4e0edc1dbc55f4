"""CPU restatement of the payload side: generator, de-generator, pattern vote.

Test infrastructure (see oracle/__init__.py).  Follows
src/offmark/generator/shuffler.py:15-25, src/offmark/degenerator/de_shuffler.py:8-22,
src/offmark/generator/grayscale.py:16-31, src/offmark/degenerator/de_grayscale.py:8-23,
the pattern vote of tests/segment_mark_detect_hls.py:144-155 (same code at
tests/mark_video_to_hls.py:273-282 and tests/detect_watermarks.py:126-137) and
the payload schemes of tests/segment_mark_detect_hls.py:42-55 and
tests/mark_video_to_hls.py:27-43.
"""
import math
from collections import Counter

import numpy as np


def permutation(length, key):
    """Index order ``RandomState(key).shuffle`` leaves ``arange(length)`` in
    (de_shuffler.py:10-11).  key 0, length 8 -> [6 2 1 7 3 0 5 4]."""
    idx = np.arange(length)
    np.random.RandomState(key).shuffle(idx)
    return idx


def generate_wm(payload, capacity, key):
    """``Shuffler.generate_wm`` (shuffler.py:15-25): shuffle a copy of the payload
    with MT19937(key), repeat to capacity.  ``wm.flat[c] == shuffled[c % len]``."""
    size = int(np.prod(capacity))
    flat = np.array(payload).copy()
    reps = int(math.ceil(size / flat.size))
    np.random.RandomState(key).shuffle(flat)
    return np.tile(flat, (reps, 1)).reshape(-1)[:size].reshape(capacity)


def generate_wm_grayscale(image, capacity, key):
    """``GrayScale.generate_wm`` (grayscale.py:16-31): threshold at 127, flatten,
    then as ``generate_wm``."""
    bits = (np.asarray(image) > 127).astype(np.uint8).flatten()
    return generate_wm(bits, capacity, key)


def position_means(wm_bits, length):
    """Mean of every ``length``-th raw bit, per payload position (de_shuffler.py:17-18)."""
    flat = np.asarray(wm_bits).flatten()
    return np.array([flat[i::length].mean() for i in range(length)])


def degenerate(wm_bits, length, key):
    """``DeShuffler.degenerate`` (de_shuffler.py:14-22) -> uint8 (length,).
    The float64 expression is kept exactly as written: means, scatter through
    the key's permutation, threshold ``0.5 * (max + min)``, strict ``>``."""
    means = position_means(wm_bits, length)
    out = np.zeros(length)
    out[permutation(length, key)] = means
    threshold = 0.5 * (np.max(out) + np.min(out))
    return (out > threshold).astype(np.uint8)


def degenerate_grayscale(wm_bits, shape, key):
    """``DeGrayScale.degenerate`` (de_grayscale.py:15-23)."""
    length = int(np.prod(shape))
    return (degenerate(wm_bits, length, key) * 255).reshape(shape)


def pattern_vote(patterns):
    """Mode of the per-frame patterns and its frequency
    (segment_mark_detect_hls.py:144-155).  ``Counter.most_common(1)`` breaks ties
    in favour of the pattern seen first.  Empty input -> (None, None) (:140-142)."""
    if len(patterns) == 0:
        return None, None
    strings = [''.join(map(str, p)) for p in patterns]
    best, count = Counter(strings).most_common(1)[0]
    return np.array([int(ch) for ch in best]), count / len(strings)


def payload_for_segment(segment_number):
    """tests/segment_mark_detect_hls.py:42-55: 8-bit big-endian segment number."""
    return np.array([int(b) for b in format(segment_number % 256, '08b')])


def payload_for_segment_copy(segment_number, copy_index):
    """tests/mark_video_to_hls.py:27-43: 4-bit segment number, then 4-bit copy index."""
    bits = format(segment_number % 16, '04b') + format(copy_index % 16, '04b')
    return np.array([int(b) for b in bits])


def segment_copy_from_pattern(pattern):
    """Inverse of ``payload_for_segment_copy`` (tests/detect_watermarks.py:145-172)."""
    bits = ''.join(str(int(b)) for b in pattern)
    return int(bits[:4], 2), int(bits[4:8], 2)
