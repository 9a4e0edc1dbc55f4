"""In-memory stand-ins for ``FileDecoder`` / ``FileEncoder``.

The reference reads and writes frames through ffmpeg pipes (src/offmark/video/frame_reader.py,
frame_writer.py); decode/encode stay outside this repository's scope, so the drivers are fed
from arrays with the same ``width`` / ``height`` / ``read()`` / ``write()`` / ``close()`` surface."""


import numpy as np


class ArrayReader:
    def __init__(self, frames):
        self._frames = iter(frames)
        first = frames[0]
        self.height, self.width = first.shape[0], first.shape[1]

    def read(self):
        return next(self._frames, None)

    def close(self):
        pass


class ArrayWriter:
    def __init__(self):
        self.frames = []

    def write(self, frame):
        self.frames.append(np.array(frame))      # a copy: the frame is consumed inside write(), like a pipe would

    def close(self):
        pass


class BatchReader:
    """Reader over one ``[N, H, W, 3]`` uint8 array (ideally the numpy view of a pinned tensor).  Besides the reference's
    ``read()`` it offers the OPTIONAL batch protocol of the B200 drivers: ``read_batch(n)`` returns a view of the next
    ``n`` frames (fewer at the end, ``None`` when exhausted) so that a whole batch goes to the GPU in one copy straight
    from the reader's memory, without the per-frame gather into a staging buffer."""

    def __init__(self, frames):
        self._frames = frames
        self._next = 0
        self.height, self.width = frames.shape[1], frames.shape[2]

    def read(self):
        batch = self.read_batch(1)
        return None if batch is None else batch[0]

    def read_batch(self, n):
        if self._next >= len(self._frames):
            return None
        out = self._frames[self._next:self._next + n]
        self._next += len(out)
        return out

    def close(self):
        pass


class BatchWriter:
    """Writer into one preallocated ``[N, H, W, 3]`` uint8 array (pinned by default).  OPTIONAL batch protocol:
    ``reserve(n, frame_shape)`` hands out the view the next ``n`` frames are to be written into (the driver downloads from
    the GPU straight into it), ``commit(n)`` publishes them; several reservations may be outstanding and are committed in
    the order they were made.  ``write(frame)`` works as in the reference (one copy)."""

    def __init__(self, n_frames, frame_shape, pinned=True):
        import torch
        self._store = torch.empty((n_frames,) + tuple(frame_shape), dtype=torch.uint8, pin_memory=pinned)
        self.array = self._store.numpy()
        self.count = 0                 # frames committed
        self._reserved = 0             # frames handed out by reserve() (>= count)

    @property
    def frames(self):
        return self.array[:self.count]

    def rewind(self):
        """Start over at frame 0 (the store is reused)."""
        self.count = self._reserved = 0

    def write(self, frame):
        if self._reserved != self.count:
            raise RuntimeError("BatchWriter.write() with reservations outstanding")
        self.array[self.count] = frame
        self.count += 1
        self._reserved = self.count

    def reserve(self, n, frame_shape):
        if tuple(frame_shape) != self.array.shape[1:] or self._reserved + n > len(self.array):
            return None
        view = self.array[self._reserved:self._reserved + n]
        self._reserved += n
        return view

    def commit(self, n):
        if self.count + n > self._reserved:
            raise RuntimeError("BatchWriter.commit() of more frames than were reserved")
        self.count += n

    def close(self):
        pass
