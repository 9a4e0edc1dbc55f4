"""Single-level float32 Haar analysis / synthesis, restating PyWavelets 1.4.1.

Test infrastructure (see oracle/__init__.py).  PyWavelets is a pinned
third-party dependency of the reference (pdm.lock:45-46) that is absent from
/root/reference and from this image.  Its call sites on the hot path are
``pywt.dwt2(plane, 'haar')`` (embed/dwt_dct_svd_encoder.py:24,
extract/dwt_dct_svd_decoder.py:19) and ``pywt.idwt2((ca, hvd), 'haar')``
(embed/dwt_dct_svd_encoder.py:26), always on even-sized float32 planes.

Published algorithm being restated (pywt/_multilevel.py ``dwtn``/``idwtn`` and
pywt/_extensions/c/convolution.template.c, float32 instantiation):

* ``dwtn`` filters axis 0 first, then axis 1; sub-band keys are built as
  ('a'|'d') per axis, and ``dwt2`` returns ``aa, (da, ad, dd)``.
* One axis, Haar, even length, any mode (no boundary taps are needed when the
  filter length is 2): ``out[k] = f[0]*x[2k+1] + f[1]*x[2k]`` accumulated in
  that order starting from 0, with dec_lo = [c, c], dec_hi = [-c, c],
  c = float32(1/sqrt(2)); every product and sum rounds to float32 (the wheels
  target baseline x86-64, so no FMA contraction).
* ``idwtn`` undoes axis 1 first, then axis 0.  One axis:
  ``out[2k] = rec_lo[0]*a[k] + rec_hi[0]*d[k]``,
  ``out[2k+1] = rec_lo[1]*a[k] + rec_hi[1]*d[k]`` with rec_lo = [c, c],
  rec_hi = [c, -c]; the approximation term is accumulated first.
"""
import numpy as np

_C = np.float32(1.0 / np.sqrt(2.0))


def _analysis_axis(x, axis):
    x = np.moveaxis(x, axis, 0)
    even, odd = x[0::2], x[1::2]
    lo = (_C * odd) + (_C * even)
    hi = (-_C * odd) + (_C * even)
    return np.moveaxis(lo, 0, axis), np.moveaxis(hi, 0, axis)


def _synthesis_axis(lo, hi, axis):
    lo = np.moveaxis(lo, axis, 0)
    hi = np.moveaxis(hi, axis, 0)
    out = np.empty((2 * lo.shape[0],) + lo.shape[1:], dtype=np.float32)
    out[0::2] = (_C * lo) + (_C * hi)
    out[1::2] = (_C * lo) + (-_C * hi)
    return np.moveaxis(out, 0, axis)


def dwt2_haar(plane):
    """``pywt.dwt2(plane, 'haar')`` for an even-sized 2-D float32 array."""
    x = np.ascontiguousarray(plane, dtype=np.float32)
    if x.ndim != 2 or x.shape[0] % 2 or x.shape[1] % 2:
        raise ValueError("oracle Haar restatement covers even-sized 2-D planes only")
    a0, d0 = _analysis_axis(x, 0)
    aa, ad = _analysis_axis(a0, 1)
    da, dd = _analysis_axis(d0, 1)
    return aa, (da, ad, dd)


def idwt2_haar(coeffs):
    """``pywt.idwt2((aa, (da, ad, dd)), 'haar')`` for float32 sub-bands."""
    aa, (da, ad, dd) = coeffs
    aa, da, ad, dd = (np.asarray(b, dtype=np.float32) for b in (aa, da, ad, dd))
    a0 = _synthesis_axis(aa, ad, 1)
    d0 = _synthesis_axis(da, dd, 1)
    return _synthesis_axis(a0, d0, 0)
