"""The two perceptual masks of the 8x8-DCT pair as public methods (dct_encoder.py:41-102, dct_decoder.py:29-89).

``encode`` / ``decode`` never materialise them (``b200wm_dct8_encode`` / ``_decode`` keep the per-block terms on the
device); these are for callers that look at the masks themselves, as the reference's public methods allow."""
import numpy as np
import torch

from b200wm import ops
from ._frames import device_of


def _block_terms(lum, device):
    """(block means f32, texture mask f32, rows, cols, numpy_in) of one float32 luminance plane ``[H, W]``."""
    numpy_in = isinstance(lum, np.ndarray)
    if numpy_in:
        if lum.ndim != 2 or lum.dtype != np.float32:
            raise ValueError("lum must be a float32 H x W array")
        lum = torch.from_numpy(np.ascontiguousarray(lum)).to(device_of(device))
    elif not isinstance(lum, torch.Tensor) or not lum.is_cuda or lum.dim() != 2 or lum.dtype != torch.float32:
        raise ValueError("lum must be a float32 H x W numpy array or CUDA tensor")
    rows, cols = lum.shape[0] // 8, lum.shape[1] // 8
    block_mean, tex, _ = ops.dct8_masks(lum)
    return block_mean[0], tex[0], rows, cols, numpy_in


class DctMasks:
    """Mixin for ``DctEncoder`` / ``DctDecoder`` (the reference carries the two methods in both classes)."""

    def luminance_mask(self, lum):
        """float64 ``[H/8, W/8]``: block DC / 8 through the brightness rule of dct_encoder.py:52-67, in float64 as
        written there (the frame mean is a float64 mean of the block means)."""
        block_mean, _, rows, cols, numpy_in = _block_terms(lum, getattr(self, "device", None))
        mask = block_mean.double()
        l_min, l_max, f_max = 90, 255, 2
        mean = max(l_min, float(mask.mean())) if mask.numel() else float("nan")
        f_ref = 1 + (mean - l_min) * (f_max - 1) / (l_max - l_min)
        bright = 1 + (mask - mean) / (l_max - mean) * (f_max - f_ref)
        one = torch.ones_like(mask)
        out = torch.where(mask > mean, bright, torch.where(mask < 15, 1.25 * one, torch.where(mask < 25, 1.125 * one, one)))
        out = out.reshape(rows, cols)
        return out.cpu().numpy() if numpy_in else out

    def texture_mask(self, lum):
        """float64 ``[H/8, W/8]``: the decision tree of dct_encoder.py:70-102 on the magnitudes of each block's DCT."""
        _, tex, rows, cols, numpy_in = _block_terms(lum, getattr(self, "device", None))
        out = tex.double().reshape(rows, cols)
        return out.cpu().numpy() if numpy_in else out
