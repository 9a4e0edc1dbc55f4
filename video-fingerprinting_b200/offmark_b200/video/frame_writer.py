"""Module name of ``offmark.video.frame_writer`` (src/offmark/video/frame_writer.py); see ``frame_reader``."""
from .memory_io import ArrayWriter      # noqa: F401

try:
    from offmark.video.frame_writer import FileEncoder      # noqa: F401
except Exception:
    class FileEncoder:
        def __init__(self, *args, **kwargs):
            raise ImportError("FileEncoder needs the reference's offmark package, ffmpeg-python and the ffmpeg binary; "
                              "use offmark_b200.video.memory_io.ArrayWriter for in-memory frames")
