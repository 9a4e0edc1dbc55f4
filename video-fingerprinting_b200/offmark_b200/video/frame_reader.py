"""Module name of ``offmark.video.frame_reader`` (src/offmark/video/frame_reader.py).

The reference's ``FileDecoder`` decodes through an ffmpeg pipe; decode stays outside this repository's scope
(DESIGN.md §7), so the name resolves to the reference's own class when the ``offmark`` package and
ffmpeg are installed, and otherwise to a loud stub.  ``ArrayReader`` is the in-memory stand-in the tests use."""
from .memory_io import ArrayReader      # noqa: F401

try:
    from offmark.video.frame_reader import FileDecoder      # noqa: F401
except Exception:      # offmark / ffmpeg-python / ffmpeg not installed
    class FileDecoder:
        def __init__(self, *args, **kwargs):
            raise ImportError("FileDecoder needs the reference's offmark package, ffmpeg-python and the ffmpeg binary; "
                              "use offmark_b200.video.memory_io.ArrayReader for in-memory frames")
