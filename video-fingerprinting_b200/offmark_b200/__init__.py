"""Drop-in mirror of the reference's plugin packages, backed by libb200wm.so.

Module paths, class names, constructor keywords and method signatures follow
``offmark`` (src/offmark/ in the reference) one to one for the modules listed below, so the
plugin imports of the reference's driver scripts (tests/mark.py:6-10, tests/detect.py:6-9,
tests/segment_mark_detect_hls.py) resolve by changing the import root from ``offmark`` to
``offmark_b200``:

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

    offmark.video.frame_reader.FileDecoder               -> offmark_b200.video.frame_reader.FileDecoder (*)
    offmark.video.frame_writer.FileEncoder               -> offmark_b200.video.frame_writer.FileEncoder (*)

(*) re-exports of the reference's ffmpeg pipes when ``offmark`` + ffmpeg are installed (decode / encode are out of
scope here), loud stubs otherwise; ``offmark_b200.video.memory_io`` has in-memory stand-ins.  NOT mirrored: the
DTCWT coders and the block / correlation shufflers that tests/test.py also imports - those imports must keep
pointing at ``offmark``.  Differences from the reference's inputs: frames must be float32 (what
video/embedder.py:34 hands to a plugin) and watermark / raw bits must be 0/1 (no soft bits).

Every ``encode`` / ``decode`` / ``degenerate`` runs CUDA kernels; there is no CPU path.
Frames may be numpy arrays (copied to the GPU and back, results written into the
caller's array like the reference's in-place ``encode``) or CUDA tensors (zero copy).
"""
