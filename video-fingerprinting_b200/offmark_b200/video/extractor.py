"""B200 drop-in for ``offmark.video.extractor`` (src/offmark/video/extractor.py)."""
import collections
import logging

import numpy as np
import torch

from b200wm import ops
from .._frames import device_of, Staging

logger = logging.getLogger(__name__)

_LANES = 2          # batches in flight in the batch protocol: one uploading while the other is read and voted


class Extractor:
    """Same frame loop as the reference (extractor.py:17-28); ``__check_frame`` (:30-34) runs the
    colour conversion, the frame plugin and the per-frame vote on the GPU and logs the pattern."""

    def __init__(self, frame_reader, frame_extractor, degenerator, device=None, batch_frames=1):
        self.frame_reader = frame_reader
        self.frame_extractor = frame_extractor
        self.degenerator = degenerator
        self.device = device
        self.batch_frames = max(1, int(batch_frames))      # optional extension: frames per kernel launch
        self.patterns = []                                 # optional extension: every pattern that was logged
        self._staging = Staging()

    def start(self):
        logger.debug('Entering start()')
        batched = (self.batch_frames > 1 and hasattr(self.frame_extractor, "decode_rgb8")
                   and hasattr(self.degenerator, "degenerate_counts"))
        if batched and hasattr(self.frame_reader, "read_batch"):
            self._run_batches()
        else:
            self._run_frames(batched)
        self.frame_reader.close()
        logger.info('Done')

    def _run_frames(self, batched):
        """The reference's loop (extractor.py:19-26): one ``read()`` per frame; batched mode gathers ``batch_frames`` of them."""
        pending = []
        while True:
            in_frame = self.frame_reader.read()
            if in_frame is None:
                logger.info('End of input stream')
                break
            if not batched:
                self._log(self.check_frame(in_frame))
                continue
            pending.append(np.ascontiguousarray(in_frame, dtype=np.uint8))
            if len(pending) == self.batch_frames:
                self._flush(pending)
        if pending:
            self._flush(pending)

    def _run_batches(self):
        """Readers with the optional batch protocol (video/memory_io.py:BatchReader) hand over ``[n, H, W, 3]`` views.
        ``_LANES`` batches are in flight on their own streams (the upload of batch k+1 overlaps the kernels of batch
        k); patterns are logged in frame order."""
        dev = device_of(self.device)
        main = torch.cuda.current_stream(dev)
        lanes = [torch.cuda.Stream(device=dev) for _ in range(_LANES)]
        inflight = collections.deque()
        k = 0
        while True:
            frames = self.frame_reader.read_batch(self.batch_frames)
            if frames is None or len(frames) == 0:
                logger.info('End of input stream')
                break
            if self.frame_extractor.scales[1] <= 0:
                while inflight:
                    self._log_batch(inflight.popleft())
                for f in frames:
                    self._log(self.check_frame(f))
                continue
            if len(inflight) == _LANES:
                self._log_batch(inflight.popleft())
            lane = lanes[k % _LANES]
            k += 1
            lane.wait_stream(main)
            with torch.cuda.stream(lane):
                host = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.uint8))
                patterns = self._launch_vote(host.to(dev, non_blocking=host.is_pinned()))
                done = torch.cuda.Event()
                done.record(lane)
            inflight.append((done, patterns))
        while inflight:
            self._log_batch(inflight.popleft())

    def _log(self, pattern):
        self.patterns.append(pattern)
        logger.info(pattern)

    def _flush(self, pending):
        """One upload, one fused extract launch and one vote launch for the whole batch; the patterns are
        logged in frame order and equal the per-frame path's."""
        dev = device_of(self.device)
        same = all(f.shape == pending[0].shape for f in pending)
        for group in ([pending] if same else [[f] for f in pending]):
            if self.frame_extractor.scales[1] <= 0:
                for f in group:
                    self._log(self.check_frame(f))
                continue
            self._vote_batch(self._staging.upload(group, dev))
        pending.clear()

    def _launch_vote(self, frames):
        """uint8 ``[n, H, W, 3]`` CUDA frames -> one fused extract launch and one vote launch on the current stream;
        returns the ``[n, payload_len]`` patterns (device)."""
        h, w = frames.shape[1], frames.shape[2]
        raw, counts = ops.dwtsvd_extract_rgb8(frames, scale=self.frame_extractor.scales[1], channel=1,
                                              payload_len=self.degenerator.payload_len)
        patterns, _ = self.degenerator.degenerate_counts(counts, h * w // 64)
        return patterns

    def _log_batch(self, batch):
        done, patterns = batch
        done.synchronize()
        fmt = getattr(self.degenerator, "format_pattern", None)
        for p in patterns.cpu().numpy():
            self._log(fmt(p) if fmt else p)

    def _vote_batch(self, frames):
        done = torch.cuda.Event()
        patterns = self._launch_vote(frames)
        done.record()
        self._log_batch((done, patterns))

    def check_frame(self, frame_rgb):
        dev = device_of(self.device)
        frame = torch.from_numpy(np.ascontiguousarray(frame_rgb, dtype=np.uint8)).to(dev)
        if hasattr(self.frame_extractor, "decode_rgb8"):     # colour conversion fused into the extract kernel
            bits = self.frame_extractor.decode_rgb8(frame)
        else:
            bits = self.frame_extractor.decode(ops.bgr8_to_yuv32(frame))
        return self.degenerator.degenerate(bits)
