"""B200 drop-in for ``offmark.video.extractor`` (src/offmark/video/extractor.py)."""
import collections
import logging

import numpy as np
import torch

from b200wm import ops
from .._frames import device_of, gathered_batches, lane_streams, Staging

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
        self._slots = [Staging() for _ in range(_LANES)]   # pinned staging, one per batch in flight

    def start(self):
        logger.debug('Entering start()')
        try:
            batched = (self.batch_frames > 1 and hasattr(self.frame_extractor, "decode_rgb8")
                       and hasattr(self.degenerator, "degenerate_counts") and self.frame_extractor.scales[1] > 0)
            if batched and hasattr(self.frame_reader, "read_batch"):
                self._run_batches()
            elif batched:
                self._run_gathered()
            else:
                self._run_frames()
        except BaseException:
            try:                                   # copies may still be in flight into the staging buffers
                torch.cuda.synchronize()
            except Exception:
                pass
            raise
        finally:
            for slot in self._slots:
                slot.release()
        self.frame_reader.close()
        logger.info('Done')

    def _run_frames(self):
        """The reference's loop (extractor.py:19-26), one frame at a time."""
        while True:
            in_frame = self.frame_reader.read()
            if in_frame is None:
                logger.info('End of input stream')
                break
            self._log(self.check_frame(in_frame))

    def _run_gathered(self):
        """The reference's ``read()`` per frame, ``batch_frames`` of them per launch (one fused extract launch and one
        vote launch), two batches in flight: the copy threads gather batch k into pinned memory while the GPU uploads
        and reads batch k-1.  The patterns are logged in frame order and equal the per-frame path's."""
        dev = device_of(self.device)
        staged = None
        for k, group in enumerate(gathered_batches(self.frame_reader, self.batch_frames)):
            fresh = self._slots[k % _LANES].stage(group)
            if staged:
                self._vote_staged(staged, dev)
            staged = fresh
        logger.info('End of input stream')
        if staged:
            self._vote_staged(staged, dev)

    def _vote_staged(self, staged, dev):
        host, futures = staged
        Staging.settle(futures)
        self._vote_batch(host.to(dev, non_blocking=True))

    def _run_batches(self):
        """Readers with the optional batch protocol (video/memory_io.py:BatchReader) hand over ``[n, H, W, 3]`` views.
        ``_LANES`` batches are in flight on their own streams (the upload of batch k+1 overlaps the kernels of batch
        k); patterns are logged in frame order."""
        dev = device_of(self.device)
        main = torch.cuda.current_stream(dev)
        lanes = lane_streams(dev, _LANES)
        inflight = collections.deque()
        k = 0
        while True:
            frames = self.frame_reader.read_batch(self.batch_frames)
            if frames is None or len(frames) == 0:
                logger.info('End of input stream')
                break
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
