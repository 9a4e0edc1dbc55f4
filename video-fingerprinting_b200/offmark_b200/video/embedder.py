"""B200 drop-in for ``offmark.video.embedder`` (src/offmark/video/embedder.py)."""
import collections
import logging

import numpy as np
import torch

from b200wm import ops
from .._frames import device_of, gathered_batches, lane_streams, Staging

logger = logging.getLogger(__name__)

_LANES = 3          # batches in flight: one being gathered / uploaded, one in the kernel, one downloaded / written


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
        self._slots = [Staging() for _ in range(_LANES)]   # pinned staging, one per batch in flight

    def start(self):
        logger.debug('Entering start()')
        try:
            batched = self.batch_frames > 1 and hasattr(self.frame_embedder, "mark_rgb8")
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
        self.frame_writer.close()
        logger.info('Done')

    def _run_frames(self):
        """The reference's loop (embedder.py:19-27), one frame at a time."""
        while True:
            in_frame = self.frame_reader.read()
            if in_frame is None:
                logger.info('End of input stream')
                break
            self.frame_writer.write(self.mark_frame(in_frame))

    # ------------------------------------------------------------------ batched mode, reference-shaped reader / writer
    def _run_gathered(self):
        """The reference's ``read()`` / ``write()`` per frame, ``batch_frames`` of them per kernel launch, as a three
        stage pipeline over ``_LANES`` staging slots: while the main thread hands batch k-2 to the writer, the copy
        threads gather batch k into pinned memory and the GPU uploads, marks and downloads batch k-1.  Frames are
        written in order and byte-identical to the per-frame path."""
        dev = device_of(self.device)
        main = torch.cuda.current_stream(dev)
        lanes = lane_streams(dev, _LANES)
        staged = flying = None                 # (slot, pinned host tensor, copy futures) / (event, marked host frames)
        for k, group in enumerate(gathered_batches(self.frame_reader, self.batch_frames)):
            slot = k % _LANES
            fresh = (slot,) + self._slots[slot].stage(group)
            launched = self._launch_staged(staged, dev, main, lanes) if staged else None
            if flying:
                self._write_out(flying)
            staged, flying = fresh, launched
        logger.info('End of input stream')
        launched = self._launch_staged(staged, dev, main, lanes) if staged else None
        for batch in (flying, launched):
            if batch:
                self._write_out(batch)

    def _launch_staged(self, staged, dev, main, lanes):
        slot, host, futures = staged
        Staging.settle(futures)
        lanes[slot].wait_stream(main)
        with torch.cuda.stream(lanes[slot]):
            marked = self._slots[slot].download_async(self.frame_embedder.mark_rgb8(host.to(dev, non_blocking=True)))
            done = torch.cuda.Event()
            done.record(lanes[slot])
        return done, marked

    def _write_out(self, batch):
        """Views of a pinned buffer that a later batch overwrites: writers consume a frame inside write(), as the
        reference's FileEncoder does (video/frame_writer.py:41-44 serialises it straight into the pipe)."""
        done, marked = batch
        done.synchronize()
        for f in marked:
            self.frame_writer.write(f)

    # ------------------------------------------------------------------ batched mode, batch protocol
    def _run_batches(self):
        """Readers with the optional batch protocol (video/memory_io.py:BatchReader) hand over ``[n, H, W, 3]`` views.
        Up to ``_LANES`` batches are in flight, each on its own stream, so that the upload of batch k+1, the kernel of
        batch k and the download of batch k-1 overlap; batches are committed to the writer in order."""
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
                self._retire(inflight.popleft())
            slot = k % _LANES
            k += 1
            lanes[slot].wait_stream(main)
            with torch.cuda.stream(lanes[slot]):
                inflight.append(self._launch_array(np.ascontiguousarray(frames, dtype=np.uint8), dev, slot, lanes[slot]))
        while inflight:
            self._retire(inflight.popleft())

    def _launch_array(self, frames, dev, slot, lane):
        """A batch that is already one ``[n, H, W, 3]`` array, on the current stream ``lane``: one copy up (asynchronous
        and at link speed when the reader's memory is pinned), one launch, one copy down - straight into the writer's
        memory when it offers ``reserve`` / ``commit``, else into this slot's pinned buffer."""
        host = torch.from_numpy(frames)
        marked = self.frame_embedder.mark_rgb8(host.to(dev, non_blocking=host.is_pinned()))
        dest = self.frame_writer.reserve(len(frames), frames.shape[1:]) if hasattr(self.frame_writer, "reserve") else None
        if dest is not None:
            torch.from_numpy(dest).copy_(marked, non_blocking=True)
            out = None
        else:
            out = self._slots[slot].download_async(marked)
        done = torch.cuda.Event()
        done.record(lane)
        return done, len(frames), out

    def _retire(self, batch):
        done, n, out = batch
        if out is None:
            done.synchronize()
            self.frame_writer.commit(n)
        else:
            self._write_out((done, out))

    def mark_frame(self, frame_rgb):
        """uint8 H x W x 3 in, uint8 H x W x 3 out (numpy)."""
        dev = device_of(self.device)
        frame = torch.from_numpy(np.ascontiguousarray(frame_rgb, dtype=np.uint8)).to(dev)
        if hasattr(self.frame_embedder, "mark_rgb8"):        # colour bracket fused into the plugin's kernel
            return self.frame_embedder.mark_rgb8(frame).cpu().numpy()
        yuv = ops.bgr8_to_yuv32(frame)
        yuv = self.frame_embedder.encode(yuv)
        return ops.yuv32_to_bgr8(yuv).cpu().numpy()
