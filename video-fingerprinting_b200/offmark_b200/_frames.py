"""Moving frames between the caller's arrays and the GPU for the drop-in plugins."""
import threading

import numpy as np
import torch

from b200wm import ops


def device_of(device):
    ops.require_cuda()
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


class FrameOnDevice:
    """One H x W x 3 float32 frame on the GPU plus the way back to where it came from."""

    def __init__(self, frame, device=None):
        self.src = frame
        if isinstance(frame, torch.Tensor):
            if not frame.is_cuda:
                raise ValueError("torch frames must live on the GPU (numpy arrays are accepted for host data)")
            if frame.dtype != torch.float32 or frame.dim() != 3 or frame.shape[2] != 3:
                raise ValueError("frame must be float32 H x W x 3")
            self.dev = frame
        else:
            if not isinstance(frame, np.ndarray) or frame.ndim != 3 or frame.shape[2] != 3:
                raise ValueError("frame must be an H x W x 3 array")
            if frame.dtype != np.float32:
                # the reference's pywt/cv2 calls promote or reject other dtypes; the drivers always pass float32
                raise ValueError("frame must be float32 (video/embedder.py:34 converts before the plugin runs)")
            self.dev = torch.from_numpy(np.ascontiguousarray(frame)).to(device_of(device))

    def write_back(self):
        """Copy the device frame into the caller's numpy array (the reference mutates its argument)."""
        if isinstance(self.src, np.ndarray):
            host = self.dev.cpu().numpy()
            self.src[...] = host
        return self.src


def gathered_batches(reader, batch_frames):
    """The reference's ``read()`` per frame, gathered into lists of up to ``batch_frames`` contiguous uint8 frames of one
    shape (a batch is one kernel launch); a frame of another shape closes the batch before it."""
    pending = []
    while True:
        frame = reader.read()
        if frame is None:
            break
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        if pending and frame.shape != pending[0].shape:
            yield pending
            pending = []
        pending.append(frame)
        if len(pending) == batch_frames:
            yield pending
            pending = []
    if pending:
        yield pending


_LANE_STREAMS = {}


def lane_streams(device, n):
    """``n`` side streams of ``device`` for batches in flight, created once per process: the caching allocator keeps one
    block pool per stream, so drivers that made their own streams on every ``start()`` paid fresh ``cudaMalloc`` calls
    for every batch buffer and left the freed blocks stranded in pools nobody used again."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    have = _LANE_STREAMS.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device=device))
    return have[:n]


_POOL = None


def _copy_pool():
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) // 2)), thread_name_prefix="b200wm-stage")
    return _POOL


_PINNED_FREE = []               # pinned buffers of finished drivers, kept for the next ones (pinning 200 MB costs ~20 ms)
_PINNED_LOCK = threading.Lock()
_PINNED_KEEP = 8


def _take_pinned(nbytes):
    with _PINNED_LOCK:
        fits = [b for b in _PINNED_FREE if b.numel() >= nbytes]
        if fits:
            best = min(fits, key=lambda b: b.numel())
            _PINNED_FREE[:] = [b for b in _PINNED_FREE if b is not best]
            return best
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def _give_pinned(buf):
    with _PINNED_LOCK:
        _PINNED_FREE.append(buf)
        _PINNED_FREE.sort(key=lambda b: b.numel())
        del _PINNED_FREE[:-_PINNED_KEEP]            # the largest ones stay


def release_staging():
    """Frees the pinned staging buffers kept between driver runs."""
    with _PINNED_LOCK:
        _PINNED_FREE.clear()


class Staging:
    """Pinned host buffers for the batched drivers: frames are gathered into ``up`` (one H2D copy per batch at
    full link speed instead of one pageable copy per frame) and results land in ``down``.  Grown on demand,
    reused between batches; ``release`` hands them to a small process-wide pool for the next driver."""

    def __init__(self):
        self.up = self.down = None

    def _fit(self, buf, shape):
        n = int(np.prod(shape))
        if buf is None or buf.numel() < n:
            if buf is not None:
                _give_pinned(buf)
            buf = _take_pinned(n)
        return buf

    def release(self):
        """Call when nothing reads the views handed out any more (the drivers do at the end of ``start()``)."""
        for buf in (self.up, self.down):
            if buf is not None:
                _give_pinned(buf)
        self.up = self.down = None

    def stage(self, frames):
        """list of equal-shape uint8 arrays -> (pinned host tensor [N, *shape], futures of the copies into it).  The
        copies run on the copy threads (numpy copies release the GIL: a few threads move a batch at several times one
        core's memcpy rate); the caller goes on and calls ``settle`` before it touches the tensor."""
        shape = (len(frames),) + tuple(frames[0].shape)
        self.up = self._fit(self.up, shape)
        host = self.up[:int(np.prod(shape))].view(shape)
        view = host.numpy()
        if len(frames) >= 4 and frames[0].nbytes >= (1 << 20):
            pool = _copy_pool()
            return host, [pool.submit(np.copyto, view[i], f, casting="unsafe") for i, f in enumerate(frames)]
        for i, f in enumerate(frames):
            np.copyto(view[i], f, casting="unsafe")
        return host, []

    @staticmethod
    def settle(futures):
        for f in futures:
            f.result()

    def upload(self, frames, device):
        """list of equal-shape uint8 arrays -> CUDA tensor [N, *shape] (async on the current stream)."""
        host, futures = self.stage(frames)
        self.settle(futures)
        return host.to(device, non_blocking=True)

    def download_async(self, tensor):
        """CUDA uint8 tensor -> numpy view of the pinned ``down`` buffer, filled asynchronously on the current stream
        (valid once that stream has passed this point, until the next download into this object)."""
        shape = tuple(tensor.shape)
        self.down = self._fit(self.down, shape)
        host = self.down[:tensor.numel()].view(shape)
        host.copy_(tensor, non_blocking=True)
        return host.numpy()

    def download(self, tensor):
        """CUDA uint8 tensor -> numpy view of the pinned ``down`` buffer (valid until the next download)."""
        host = self.download_async(tensor)
        torch.cuda.current_stream().synchronize()
        return host
