"""Haar-only stand-in for PyWavelets, used ONLY by oracle/make_golden.py.

PyWavelets 1.4.1 is a dependency of the reference (pdm.lock:45-46) that is not
installed in this image and cannot be fetched.  The reference modules
embed/dwt_dct_svd_encoder.py and extract/dwt_dct_svd_decoder.py do
``import pywt`` at module top, so to run them *verbatim* in this container the
golden-vector generator puts this directory on ``sys.path``.  It forwards to
the restatement in ``oracle/haar.py``; nothing else in the repo imports it.
"""
from oracle.haar import dwt2_haar, idwt2_haar

__version__ = "1.4.1-haar-standin"


def _check(wavelet):
    if wavelet != "haar":
        raise NotImplementedError("stand-in implements the 'haar' wavelet only")


def dwt2(data, wavelet, mode="symmetric", axes=(-2, -1)):
    _check(wavelet)
    return dwt2_haar(data)


def idwt2(coeffs, wavelet, mode="symmetric", axes=(-2, -1)):
    _check(wavelet)
    return idwt2_haar(coeffs)
