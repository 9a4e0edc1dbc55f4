"""Build libb200wm.so in-tree with nvcc for sm_100a.

    python video-fingerprinting_b200/build.py [--force] [--verbose]

The library has no Python or torch dependency: it is a plain C-ABI shared object
(include/b200wm.h) that the ctypes binding in ``b200wm/_lib.py`` loads.  nvcc
cross-compiles without a GPU, so this runs in the build container and the ``.so``
travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libb200wm.so")
SOURCES = ["api.cu", "dwtsvd.cu", "dwtsvd_tma.cu", "dwtsvd_copies.cu", "dct8.cu", "vote.cu", "bracket.cu", "fused_rgb.cu", "attacks.cu", "pipeline.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "b200wm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-fmad=false",     # no implicit contraction: every FMA is written as fmaf, so all kernel families round identically
    "-cudart", "static",
    "--threads", "0",  # one compilation per source file in parallel
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libb200wm.so cannot be built")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into lib/libb200wm.so.  Returns the path."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("B200WM_NVCC_EXTRA", "").split()      # tuning experiments only (e.g. -DB200WM_EXTRACT_STAGES=2)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB_PATH] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libb200wm.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
