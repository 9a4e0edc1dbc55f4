"""ctypes binding of libb200wm.so (the C ABI declared in include/b200wm.h).

This is the whole reference-side FFI: a maintainer of offmark-py who wanted the
B200 path behind their own plugin classes would copy this file.
There is no fallback of any kind: if the shared library is missing the import
of this module raises, and every non-zero status from the library raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libb200wm.so")

OK, ERR_INVALID, ERR_SHORT_WM, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
U8, F32 = 0, 1


class Plane(C.Structure):
    """``b200wm_plane`` (include/b200wm.h)."""
    _fields_ = [("dtype", C.c_int32), ("n_frames", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("pitch_bytes", C.c_int64), ("frame_stride_bytes", C.c_int64),
                ("elem_stride", C.c_int32), ("reserved", C.c_int32)]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_PP = C.POINTER(Plane)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
PROTOTYPES = {
    "b200wm_version": (C.c_int, []),
    "b200wm_strerror": (C.c_char_p, [C.c_int]),
    "b200wm_last_cuda_error": (C.c_char_p, []),
    "b200wm_device_ok": (C.c_int, []),
    "b200wm_kernel_launches": (C.c_int, []),
    "b200wm_set_path": (C.c_int, [C.c_int]),
    "b200wm_get_path": (C.c_int, []),
    "b200wm_block_num": (_i64, [C.c_int, C.c_int]),
    "b200wm_tile_count": (_i64, [C.c_int, C.c_int]),
    "b200wm_words_per_frame": (_i32, [C.c_int, C.c_int]),
    "b200wm_dwtsvd_embed": (C.c_int, [_vp, _vp, _PP, _vp, _i32, _i32, _i64, _vp, _f32, _vp]),
    "b200wm_dwtsvd_embed_copies": (C.c_int, [_vp, _PP, _vp, _i64, _i32, _vp, _i32, _i32, _i64, _vp, _f32, _vp]),
    "b200wm_dwtsvd_extract": (C.c_int, [_vp, _PP, _f32, _vp, _i32, _i32, _vp, _vp]),
    "b200wm_dwtsvd_sigma": (C.c_int, [_vp, _PP, _vp, _vp]),
    "b200wm_dwtsvd_sigma_dct": (C.c_int, [_vp, _PP, _vp, _vp]),
    "b200wm_dct8_masks": (C.c_int, [_vp, _PP, _vp, _vp, _vp, _vp]),
    "b200wm_dct8_embed": (C.c_int, [_vp, _vp, _PP, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _f32, _vp]),
    "b200wm_dct8_extract": (C.c_int, [_vp, _PP, _vp, _vp, _vp, _f32, _vp, _i32, _i32, _vp, _vp]),
    "b200wm_dct8_encode": (C.c_int, [_vp, _PP, _vp, _vp, _PP, _vp, _vp, _i32, _i32, _i64, _vp, _f32, _vp]),
    "b200wm_dct8_decode": (C.c_int, [_vp, _PP, _vp, _PP, _vp, _f32, _vp, _i32, _i32, _vp, _vp]),
    "b200wm_vote_counts": (C.c_int, [_vp, _i32, _i32, _i64, _i32, _vp, _vp]),
    "b200wm_vote_finish": (C.c_int, [_vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp]),
    "b200wm_pattern_hist": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "b200wm_pattern_hist_publish": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                              C.c_uint32, _vp, _vp, _vp]),
    "b200wm_vote_exchange_wait": (C.c_int, [_vp, _i64, _i32, _i32, C.c_uint32, _vp, _vp]),
    "b200wm_vote_state_reset": (C.c_int, [_vp, _i64, _i64, _vp]),
    "b200wm_bgr8_to_yuv32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "b200wm_yuv32_to_bgr8": (C.c_int, [_vp, _vp, _i64, _vp]),
    "b200wm_dwtsvd_embed_rgb8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _i32, _i32, _i64, _vp, _vp]),
    "b200wm_dwtsvd_extract_rgb8": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _i64, _i32, _f32, _vp, _i32, _i32, _vp, _vp]),
    "b200wm_dwtsvd_mark_host": (C.c_int, [_vp, _vp, _PP, _vp, _i32, _i32, _i64, _vp, _f32, _i32]),
    "b200wm_dwtsvd_detect_host": (C.c_int, [_vp, _PP, _f32, _i32, _vp, _vp, _vp, _vp, _i32]),
    "b200wm_dwtsvd_mark_verify_host": (C.c_int, [_vp, _vp, _PP, _vp, _i32, _i32, _i64, _vp, _f32, _i32, _vp, _vp, _vp, _vp, _i32]),
    "b200wm_host_scratch_release": (C.c_int, []),
    "b200wm_attack_jpeg_requant": (C.c_int, [_vp, _vp, _PP, _i32, _vp]),
    "b200wm_attack_add_noise": (C.c_int, [_vp, _vp, _PP, _vp, _vp]),
    "b200wm_attack_resize": (C.c_int, [_vp, _PP, _vp, _PP, _i32, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python video-fingerprinting_b200/build.py` "
            "(there is no CPU or PyTorch fallback for the watermark kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


class B200wmError(RuntimeError):
    pass


def check(status):
    """Map a b200wm_status to the exception the reference's own code would raise."""
    if status == OK:
        return
    text = lib.b200wm_strerror(status).decode()
    if status == ERR_SHORT_WM:
        raise IndexError(text)                      # embed/dwt_dct_svd_encoder.py:36 raises IndexError
    if status in (ERR_INVALID, ERR_UNSUPPORTED):
        raise ValueError(text)
    detail = lib.b200wm_last_cuda_error().decode()
    raise B200wmError(f"{text}: {detail}" if detail else text)
