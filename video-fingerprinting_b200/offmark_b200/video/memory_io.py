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
