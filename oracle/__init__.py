"""CPU oracle for the offmark-py watermark hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product path
(``video-fingerprinting_b200/``) never imports this package and raises if
the CUDA library is missing.

What it restates (all citations relative to /root/reference):

* ``haar``        PyWavelets 1.4.1 (pdm.lock:45-46, not vendored) single-level
                  float32 Haar ``dwt2`` / ``idwt2`` as called at
                  src/offmark/embed/dwt_dct_svd_encoder.py:24,26 and
                  src/offmark/extract/dwt_dct_svd_decoder.py:19.
* ``dwt_dct_svd`` src/offmark/embed/dwt_dct_svd_encoder.py:19-45 and
                  src/offmark/extract/dwt_dct_svd_decoder.py:12-37.
* ``dct8``        src/offmark/embed/dct_encoder.py:18-102 and
                  src/offmark/extract/dct_decoder.py:10-89.
* ``payload``     src/offmark/generator/shuffler.py:15-25,
                  src/offmark/degenerator/de_shuffler.py:8-22, the pattern
                  vote of tests/segment_mark_detect_hls.py:144-155 and the
                  payload schemes of tests/segment_mark_detect_hls.py:42-55 /
                  tests/mark_video_to_hls.py:27-43.
* ``bracket``     src/offmark/video/embedder.py:33-39 and
                  src/offmark/video/extractor.py:30-34.

Parity pinning: the reference holds no golden vectors or asserting tests for
this path, and PyWavelets is not installable here, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``make_golden.py``
imports the reference modules verbatim from /root/reference/src (with the
Haar-only ``pywt`` stand-in in ``oracle/_shim`` that forwards to ``haar``)
and both checks oracle == reference bit for bit and writes the fixtures under
``tests/golden/``.  Residual, stated openly: the stand-in reproduces
PyWavelets' float32 operation order from its published C source, but has not
been compared with a PyWavelets binary.
"""
