"""B200 drop-in for ``offmark.video.embedder`` (src/offmark/video/embedder.py)."""
import collections
import logging

import numpy as np
import torch

from b200wm import ops
from .._frames import device_of, Staging

logger = logging.getLogger(__name__)

_LANES = 3          # batches in flight in the batch protocol: one uploading, one in the kernel, one downloading


class Embedder:
    """Same frame loop as the reference (embedder.py:17-31).  ``__mark_frame`` (:33-39) keeps the
    frame on the GPU from the uint8 upload to the uint8 download: colour conversion, the frame
    plugin's kernels, conversion back, clip and round-half-even all run there."""

    def __init__(self, frame_reader, frame_embedder, frame_writer, device=None, batch_frames=1):
        self.frame_reader = frame_reader
        self.frame_writer = frame_writer
        self.frame_embedder = frame_embedder
        self.device = device
        self.batch_frames = max(1, int(batch_frames))      # optional extension: frames per kernel launch
        self._staging = Staging()

    def start(self):
        logger.debug('Entering start()')
        batched = self.batch_frames > 1 and hasattr(self.frame_embedder, "mark_rgb8")
        if batched and hasattr(self.frame_reader, "read_batch"):
            self._run_batches()
        else:
            self._run_frames(batched)
        self.frame_reader.close()
        self.frame_writer.close()
        logger.info('Done')

    def _run_frames(self, batched):
        """The reference's loop (embedder.py:19-27): one ``read()`` per frame; batched mode gathers ``batch_frames`` of them."""
        pending = []
        while True:
            in_frame = self.frame_reader.read()
            if in_frame is None:
                logger.info('End of input stream')
                break
            if not batched:
                self.frame_writer.write(self.mark_frame(in_frame))
                continue
            pending.append(np.ascontiguousarray(in_frame, dtype=np.uint8))
            if len(pending) == self.batch_frames:
                self._flush(pending)
        if pending:
            self._flush(pending)

    def _run_batches(self):
        """Readers with the optional batch protocol (video/memory_io.py:BatchReader) hand over ``[n, H, W, 3]`` views.
        Up to ``_LANES`` batches are in flight, each on its own stream, so that the upload of batch k+1, the kernel of
        batch k and the download of batch k-1 overlap; batches are committed to the writer in order."""
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
            if len(inflight) == _LANES:
                self._retire(inflight.popleft())
            lane = lanes[k % _LANES]
            k += 1
            lane.wait_stream(main)
            with torch.cuda.stream(lane):
                inflight.append(self._launch_array(np.ascontiguousarray(frames, dtype=np.uint8), dev, lane))
        while inflight:
            self._retire(inflight.popleft())

    def _flush(self, pending):
        """One upload, one fused kernel launch and one download for the whole batch; frames are written
        in order, byte-identical to the per-frame path."""
        dev = device_of(self.device)
        same = all(f.shape == pending[0].shape for f in pending)
        groups = [pending] if same else [[f] for f in pending]
        for group in groups:
            frames = self._staging.upload(group, dev)
            # views of a pinned buffer that the next batch overwrites: writers consume a frame inside write(), as
            # the reference's FileEncoder does (video/frame_writer.py:41-44 serialises it straight into the pipe)
            marked = self._staging.download(self.frame_embedder.mark_rgb8(frames))
            for f in marked:
                self.frame_writer.write(f)
        pending.clear()

    def _launch_array(self, frames, dev, lane):
        """A batch that is already one ``[n, H, W, 3]`` array, on the current stream ``lane``: one copy up (asynchronous
        and at link speed when the reader's memory is pinned), one launch, one copy down - straight into the writer's
        memory when it offers ``reserve`` / ``commit``.  Returns what ``_retire`` needs to hand the batch over."""
        host = torch.from_numpy(frames)
        marked = self.frame_embedder.mark_rgb8(host.to(dev, non_blocking=host.is_pinned()))
        dest = self.frame_writer.reserve(len(frames), frames.shape[1:]) if hasattr(self.frame_writer, "reserve") else None
        if dest is not None:
            torch.from_numpy(dest).copy_(marked, non_blocking=True)
        done = torch.cuda.Event()
        done.record(lane)
        return done, len(frames), dest, marked

    def _retire(self, batch):
        done, n, dest, marked = batch
        done.synchronize()
        if dest is not None:
            self.frame_writer.commit(n)
            return
        for f in self._staging.download(marked):
            self.frame_writer.write(f)

    def mark_frame(self, frame_rgb):
        """uint8 H x W x 3 in, uint8 H x W x 3 out (numpy)."""
        dev = device_of(self.device)
        frame = torch.from_numpy(np.ascontiguousarray(frame_rgb, dtype=np.uint8)).to(dev)
        if hasattr(self.frame_embedder, "mark_rgb8"):        # colour bracket fused into the plugin's kernel
            return self.frame_embedder.mark_rgb8(frame).cpu().numpy()
        yuv = ops.bgr8_to_yuv32(frame)
        yuv = self.frame_embedder.encode(yuv)
        return ops.yuv32_to_bgr8(yuv).cpu().numpy()
