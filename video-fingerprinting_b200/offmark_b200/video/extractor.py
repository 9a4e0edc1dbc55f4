"""B200 drop-in for ``offmark.video.extractor`` (src/offmark/video/extractor.py)."""
import logging

import numpy as np
import torch

from b200wm import ops
from .._frames import device_of

logger = logging.getLogger(__name__)


class Extractor:
    """Same frame loop as the reference (extractor.py:17-28); ``__check_frame`` (:30-34) runs the
    colour conversion, the frame plugin and the per-frame vote on the GPU and logs the pattern."""

    def __init__(self, frame_reader, frame_extractor, degenerator, device=None):
        self.frame_reader = frame_reader
        self.frame_extractor = frame_extractor
        self.degenerator = degenerator
        self.device = device

    def start(self):
        logger.debug('Entering start()')
        while True:
            in_frame = self.frame_reader.read()
            if in_frame is None:
                logger.info('End of input stream')
                break
            logger.info(self.check_frame(in_frame))
        self.frame_reader.close()
        logger.info('Done')

    def check_frame(self, frame_rgb):
        dev = device_of(self.device)
        frame = torch.from_numpy(np.ascontiguousarray(frame_rgb, dtype=np.uint8)).to(dev)
        if hasattr(self.frame_extractor, "decode_rgb8"):     # colour conversion fused into the extract kernel
            bits = self.frame_extractor.decode_rgb8(frame)
        else:
            bits = self.frame_extractor.decode(ops.bgr8_to_yuv32(frame))
        return self.degenerator.degenerate(bits)
