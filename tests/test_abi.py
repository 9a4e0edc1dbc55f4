"""CPU: the C-ABI library loads and exports exactly what include/b200wm.h declares.
No kernel is launched here (no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b200wm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"B200WM_API\s+[\w\s\*]+?\b(b200wm_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from b200wm import _lib
    names = _header_functions()
    assert len(names) >= 27
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in b200wm.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototype table and header disagree"


def test_version_and_strerror():
    from b200wm import _lib
    assert _lib.lib.b200wm_version() == 1
    assert _lib.lib.b200wm_strerror(0) == b"ok"
    assert b"shorter" in _lib.lib.b200wm_strerror(_lib.ERR_SHORT_WM)


def test_status_mapping_follows_reference_exceptions():
    from b200wm import _lib
    with pytest.raises(IndexError):          # embed/dwt_dct_svd_encoder.py:36 raises IndexError
        _lib.check(_lib.ERR_SHORT_WM)
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_INVALID)
    with pytest.raises(_lib.B200wmError):
        _lib.check(_lib.ERR_CUDA)


@pytest.mark.parametrize("h,w", [(1080, 1920), (2160, 3840), (240, 320), (1082, 1922), (37, 53), (7, 9), (8, 8), (4, 4)])
def test_geometry_matches_reference_rules(h, w):
    from b200wm import ops
    from oracle import dwt_dct_svd as svd
    block_num, tiles, words = ops.geometry(h, w)
    assert block_num == svd.wm_capacity((h, w, 3))[1] == h * w // 64
    nr, nc = svd.block_grid(h, w)
    assert tiles == nr * nc
    assert words == (block_num + 31) // 32
    assert tiles <= block_num


def test_plane_struct_layout():
    from b200wm._lib import Plane
    assert ctypes.sizeof(Plane) == 40
    assert Plane.pitch_bytes.offset == 16 and Plane.elem_stride.offset == 32


def test_pack_unpack_bits_roundtrip():
    import torch
    from b200wm import ops
    rng = np.random.RandomState(3)
    for n in (1, 31, 32, 33, 1200, 32400):
        bits = rng.randint(0, 2, (3, n))
        packed, length = ops.pack_bits(bits)
        assert length == n and packed.dtype == torch.int32 and packed.shape == (3, max(1, (n + 31) // 32))
        assert np.array_equal(ops.unpack_bits(packed, n), bits)
        words = packed.numpy().view(np.uint32)
        for c in (0, n // 2, n - 1):
            assert (words[1, c >> 5] >> (c & 31)) & 1 == bits[1, c]
    with pytest.raises(ValueError):
        ops.pack_bits(np.array([0, 2, 1]))


def _build_c_demo(tmp_path):
    import shutil
    import subprocess
    from b200wm import _lib
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if not gcc or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA headers are not available")
    exe = str(tmp_path / "c_abi_demo")
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L" + lib_dir, "-lb200wm",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + lib_dir]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def test_plain_c_caller_compiles_and_links(tmp_path):
    """The boundary is usable from C with nothing but include/b200wm.h and the shared library."""
    assert os.path.exists(_build_c_demo(tmp_path))


@pytest.mark.gpu
def test_plain_c_caller_recovers_the_payload(tmp_path):
    import subprocess
    proc = subprocess.run([_build_c_demo(tmp_path)], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "recovered payload: 0 1 1 0 0 1 0 1" in proc.stdout and "matches" in proc.stdout


def test_ctypes_prototypes_have_the_headers_parameter_counts():
    """Every binding in b200wm/_lib.py declares as many arguments as the C prototype in include/b200wm.h."""
    from b200wm import _lib
    text = open(os.path.join(ROOT, "include", "b200wm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for m in re.finditer(r"B200WM_API\s+[\w\s\*]+?\b(b200wm_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert name in _lib.PROTOTYPES, name
        assert len(_lib.PROTOTYPES[name][1]) == n, f"{name}: header has {n} parameters, binding {len(_lib.PROTOTYPES[name][1])}"
