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


class RawBits(np.ndarray):
    """float64 (1, N) array of raw bits, as the reference's ``decode`` returns, that also keeps
    the packed bits and per-position counts the extract kernel left on the GPU so that
    ``DeShuffler.degenerate`` can finish the vote there."""

    def __new__(cls, values, packed=None, block_num=None):
        obj = np.asarray(values, dtype=np.float64).view(cls)
        obj.packed = packed
        obj.block_num = block_num
        return obj

    def __array_finalize__(self, obj):
        self.packed = getattr(obj, "packed", None)
        self.block_num = getattr(obj, "block_num", None)
