"""Moving frames between the caller's arrays and the GPU for the drop-in plugins."""
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


_POOL = None


def _copy_pool():
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) // 2)), thread_name_prefix="b200wm-stage")
    return _POOL


class Staging:
    """Pinned host buffers for the batched drivers: frames are gathered into ``up`` (one H2D copy per batch at
    full link speed instead of one pageable copy per frame) and results land in ``down``.  Grown on demand,
    reused between batches."""

    def __init__(self):
        self.up = self.down = None

    def _fit(self, buf, shape):
        n = int(np.prod(shape))
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        return buf

    def upload(self, frames, device):
        """list of equal-shape uint8 arrays -> CUDA tensor [N, *shape] (async on the current stream)."""
        shape = (len(frames),) + tuple(frames[0].shape)
        self.up = self._fit(self.up, shape)
        host = self.up[:int(np.prod(shape))].view(shape)
        view = host.numpy()
        if len(frames) >= 4 and frames[0].nbytes >= (1 << 20):
            # numpy copies release the GIL: a few threads move the batch at several times one core's memcpy rate
            list(_copy_pool().map(lambda i: np.copyto(view[i], frames[i], casting="unsafe"), range(len(frames))))
        else:
            for i, f in enumerate(frames):
                np.copyto(view[i], f, casting="unsafe")
        return host.to(device, non_blocking=True)

    def download(self, tensor):
        """CUDA uint8 tensor -> numpy view of the pinned ``down`` buffer (valid until the next download)."""
        shape = tuple(tensor.shape)
        self.down = self._fit(self.down, shape)
        host = self.down[:tensor.numel()].view(shape)
        host.copy_(tensor, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()
