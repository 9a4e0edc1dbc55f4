"""B200 drop-in for ``offmark.video.embedder`` (src/offmark/video/embedder.py)."""
import logging

import numpy as np
import torch

from b200wm import ops
from .._frames import device_of

logger = logging.getLogger(__name__)


class Embedder:
    """Same frame loop as the reference (embedder.py:17-31).  ``__mark_frame`` (:33-39) keeps the
    frame on the GPU from the uint8 upload to the uint8 download: colour conversion, the frame
    plugin's kernels, conversion back, clip and round-half-even all run there."""

    def __init__(self, frame_reader, frame_embedder, frame_writer, device=None):
        self.frame_reader = frame_reader
        self.frame_writer = frame_writer
        self.frame_embedder = frame_embedder
        self.device = device

    def start(self):
        logger.debug('Entering start()')
        while True:
            in_frame = self.frame_reader.read()
            if in_frame is None:
                logger.info('End of input stream')
                break
            self.frame_writer.write(self.mark_frame(in_frame))
        self.frame_reader.close()
        self.frame_writer.close()
        logger.info('Done')

    def mark_frame(self, frame_rgb):
        """uint8 H x W x 3 in, uint8 H x W x 3 out (numpy)."""
        dev = device_of(self.device)
        frame = torch.from_numpy(np.ascontiguousarray(frame_rgb, dtype=np.uint8)).to(dev)
        if hasattr(self.frame_embedder, "mark_rgb8"):        # colour bracket fused into the plugin's kernel
            return self.frame_embedder.mark_rgb8(frame).cpu().numpy()
        yuv = ops.bgr8_to_yuv32(frame)
        yuv = self.frame_embedder.encode(yuv)
        return ops.yuv32_to_bgr8(yuv).cpu().numpy()
