"""CPU restatement of the colour-space bracket around the frame plugins.

Test infrastructure (see oracle/__init__.py).  Follows
src/offmark/video/embedder.py:33-39 (``__mark_frame``) and
src/offmark/video/extractor.py:30-34 (``__check_frame``).  ``cv2.cvtColor`` is
OpenCV (pinned 4.6.0.66 in pdm.lock:31-32, 4.13 in this image); on float32
input it uses the 0.5 chroma offset:

    Y = .114 c0 + .587 c1 + .299 c2,  U = .492 (c0 - Y) + .5,  V = .877 (c2 - Y) + .5
    c0 = Y + 2.032 (U - .5),  c1 = Y - .395 (U - .5) - .581 (V - .5),  c2 = Y + 1.14 (V - .5)

``FileDecoder`` hands over rgb24 frames (src/offmark/video/frame_reader.py:59-63)
which the bracket treats as BGR, so c0 is really R; the arithmetic does not care.
"""
import numpy as np
import cv2


def to_yuv(frame_u8):
    """embedder.py:34 / extractor.py:31."""
    return cv2.cvtColor(frame_u8.astype(np.float32), cv2.COLOR_BGR2YUV)


def from_yuv(yuv):
    """embedder.py:36-38: back to BGR, clip, round half to even, uint8."""
    bgr = cv2.cvtColor(yuv, cv2.COLOR_YUV2BGR)
    return np.around(np.clip(bgr, a_min=0, a_max=255)).astype(np.uint8)


def mark_frame(frame_u8, encode):
    """``Embedder.__mark_frame`` with ``encode(yuv) -> yuv`` as the frame plugin."""
    return from_yuv(encode(to_yuv(frame_u8)))


def check_frame(frame_u8, decode, degenerate):
    """``Extractor.__check_frame``: raw bits then the per-frame vote."""
    return degenerate(decode(to_yuv(frame_u8)))
