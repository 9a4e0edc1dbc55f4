"""Drop-in mirror of the reference's plugin packages, backed by libb200wm.so.

Module paths, class names, constructor keywords and method signatures follow
``offmark`` (src/offmark/ in the reference) one to one, so the reference's driver
scripts (tests/mark.py, tests/detect.py, tests/test.py,
tests/segment_mark_detect_hls.py) run against the B200 path by changing only the
import root from ``offmark`` to ``offmark_b200``:

    offmark.embed.dwt_dct_svd_encoder.DwtDctSvdEncoder   -> offmark_b200.embed.dwt_dct_svd_encoder.DwtDctSvdEncoder
    offmark.extract.dwt_dct_svd_decoder.DwtDctSvdDecoder -> offmark_b200.extract.dwt_dct_svd_decoder.DwtDctSvdDecoder
    offmark.embed.dct_encoder.DctEncoder                 -> offmark_b200.embed.dct_encoder.DctEncoder
    offmark.extract.dct_decoder.DctDecoder               -> offmark_b200.extract.dct_decoder.DctDecoder
    offmark.generator.shuffler.Shuffler                  -> offmark_b200.generator.shuffler.Shuffler
    offmark.generator.grayscale.GrayScale                -> offmark_b200.generator.grayscale.GrayScale
    offmark.degenerator.de_shuffler.DeShuffler           -> offmark_b200.degenerator.de_shuffler.DeShuffler
    offmark.degenerator.de_grayscale.DeGrayScale         -> offmark_b200.degenerator.de_grayscale.DeGrayScale
    offmark.video.embedder.Embedder                      -> offmark_b200.video.embedder.Embedder
    offmark.video.extractor.Extractor                    -> offmark_b200.video.extractor.Extractor

Every ``encode`` / ``decode`` / ``degenerate`` runs CUDA kernels; there is no CPU path.
Frames may be numpy arrays (copied to the GPU and back, results written into the
caller's array like the reference's in-place ``encode``) or CUDA tensors (zero copy).
"""
