"""CPU: the product path has no CPU fallback and never touches the oracle."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import PKG

cpu_only = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not pat.search(text), f"{f} reaches into oracle/ or the reference"


@cpu_only
def test_ops_raise_without_cuda():
    from b200wm import ops, B200wmError
    with pytest.raises(B200wmError):
        ops.require_cuda()


@cpu_only
def test_plugins_raise_without_cuda():
    from b200wm import B200wmError
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.degenerator.de_shuffler import DeShuffler
    yuv = np.zeros((16, 16, 3), dtype=np.float32)
    enc = DwtDctSvdEncoder()
    enc.read_wm(np.zeros((1, 4), dtype=np.int64))
    with pytest.raises(B200wmError):
        enc.encode(yuv)
    with pytest.raises(B200wmError):
        DwtDctSvdDecoder().decode(yuv)
    with pytest.raises(B200wmError):
        DeShuffler(key=0).set_shape((8,)).degenerate(np.zeros((1, 64)))


def test_constructor_surface_matches_reference():
    from offmark_b200.embed.dwt_dct_svd_encoder import DwtDctSvdEncoder
    from offmark_b200.extract.dwt_dct_svd_decoder import DwtDctSvdDecoder
    from offmark_b200.embed.dct_encoder import DctEncoder
    from offmark_b200.extract.dct_decoder import DctDecoder
    e = DwtDctSvdEncoder()
    assert (e.key, e.scales, e.blk) == (None, [0, 15, 0], 4)
    assert e.wm_capacity((1080, 1920, 3)) == (1, 32400)
    d = DwtDctSvdDecoder(key=3, scales=[0, 20, 0])
    assert (d.key, d.scales, d.blk) == (3, [0, 20, 0], 4)
    assert (DctEncoder().alpha, DctDecoder(alpha=12).alpha) == (20, 12)
    assert DctEncoder().wm_capacity((240, 320, 3)) == (1, 1200)
    with pytest.raises(ValueError):
        DwtDctSvdEncoder(blk=8)
