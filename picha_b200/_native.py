"""ctypes binding of include/picha_b200.h (the same C-ABI picha's Node addon would bind).

The library is built in-tree (picha_b200/libpicha_b200.so) by ``picha_b200.build`` /
``__graft_entry__.build()``.  There is no CPU fallback anywhere: if the library is missing
the import of this module fails, and if no CUDA device is usable every compute entry point
returns PICHA_B200_ERR_NO_DEVICE, which the host layer raises as an Error.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PICHA_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("PICHA_B200_LIB") or os.path.join(_HERE, "libpicha_b200.so")

PIXELS = ["rgb", "rgba", "grey", "greya", "r16", "r16g16", "r16g16b16", "r16g16b16a16"]
FILTERS = ["cubic", "lanczos", "catmulrom", "mitchel", "box", "triangle"]

OK = 0
ERR_INVALID_IMAGE = -1
ERR_INVALID_DIMENSIONS = -2
ERR_INVALID_FILTER = -3
ERR_INVALID_FILTER_WIDTH = -4
ERR_INVALID_PIXEL = -5
ERR_FORMAT_MISMATCH = -6
ERR_SIZE_MISMATCH = -7
ERR_NO_DEVICE = -8
ERR_CUDA = -9
ERR_NOMEM = -10
ERR_UNSUPPORTED = -11
ERR_INVALID_ARGUMENT = -12

EXACT = 1
FORCE_FAST = 2


class CImage(ctypes.Structure):
    """struct picha_b200_image == NativeImage (src/picha.h:202-218)."""
    _fields_ = [("data", ctypes.c_void_p), ("stride", ctypes.c_int32), ("width", ctypes.c_int32),
                ("height", ctypes.c_int32), ("pixel", ctypes.c_int32)]


_IMG_P = ctypes.POINTER(CImage)
_ip = ctypes.POINTER(ctypes.c_int)
_fp = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes): every symbol include/picha_b200.h declares
SIGNATURES = {
    "picha_b200_version": (ctypes.c_int, []),
    "picha_b200_device_count": (ctypes.c_int, []),
    "picha_b200_init": (ctypes.c_int, [ctypes.c_int]),
    "picha_b200_shutdown": (None, []),
    "picha_b200_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "picha_b200_last_error": (ctypes.c_char_p, []),
    "picha_b200_launch_count": (ctypes.c_uint64, []),
    "picha_b200_last_resize_kernel": (ctypes.c_int, []),
    "picha_b200_pixel_bytes": (ctypes.c_int, [ctypes.c_int]),
    "picha_b200_pixel_channels": (ctypes.c_int, [ctypes.c_int]),
    "picha_b200_row_stride": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "picha_b200_resolve_resize_options": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                         _ip, _fp]),
    "picha_b200_resolve_color_settings": (None, [ctypes.c_double, ctypes.c_double, ctypes.c_double, _fp]),
    "picha_b200_resize": (ctypes.c_int, [_IMG_P, _IMG_P, ctypes.c_int, ctypes.c_float]),
    "picha_b200_resize_ex": (ctypes.c_int, [_IMG_P, _IMG_P, ctypes.c_int, ctypes.c_float, ctypes.c_uint]),
    "picha_b200_color_convert": (ctypes.c_int, [_IMG_P, _IMG_P, ctypes.c_float, ctypes.c_float, ctypes.c_float]),
    "picha_b200_resize_batch": (ctypes.c_int, [ctypes.c_int, _IMG_P, _IMG_P, ctypes.c_int, ctypes.c_float,
                                               ctypes.c_uint, ctypes.c_int]),
    "picha_b200_cmyk_to_rgb": (ctypes.c_int, [_IMG_P, _IMG_P]),
    "picha_b200_cmyk_to_rgb_device": (ctypes.c_int, [ctypes.c_int, _IMG_P, ctypes.c_int64, _IMG_P, ctypes.c_int64, ctypes.c_void_p]),
    "picha_b200_color_convert_batch": (ctypes.c_int, [ctypes.c_int, _IMG_P, _IMG_P, ctypes.c_float, ctypes.c_float,
                                                      ctypes.c_float, ctypes.c_int]),
    "picha_b200_resize_convert": (ctypes.c_int, [_IMG_P, _IMG_P, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                 ctypes.c_float, ctypes.c_uint]),
    "picha_b200_resize_convert_batch": (ctypes.c_int, [ctypes.c_int, _IMG_P, _IMG_P, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                       ctypes.c_float, ctypes.c_float, ctypes.c_uint, ctypes.c_int]),
    "picha_b200_resize_convert_device": (ctypes.c_int, [ctypes.c_int, _IMG_P, ctypes.c_int64, _IMG_P, ctypes.c_int64, ctypes.c_int,
                                                        ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                        ctypes.c_uint, ctypes.c_void_p]),
    "picha_b200_shard_range": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _ip, _ip]),
    "picha_b200_plan_batch": (ctypes.c_int, [ctypes.c_int, _IMG_P, _IMG_P, ctypes.c_int, _ip, _ip, ctypes.c_int]),
    "picha_b200_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "picha_b200_host_free": (None, [ctypes.c_void_p]),
    "picha_b200_resize_device": (ctypes.c_int, [ctypes.c_int, _IMG_P, ctypes.c_int64, _IMG_P, ctypes.c_int64,
                                                ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.c_void_p]),
    "picha_b200_color_convert_device": (ctypes.c_int, [ctypes.c_int, _IMG_P, ctypes.c_int64, _IMG_P, ctypes.c_int64,
                                                       ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                       ctypes.c_void_p]),
    "picha_b200_synthetic_fill_device": (ctypes.c_int, [ctypes.c_int, _IMG_P, ctypes.c_int64, ctypes.c_uint64,
                                                        ctypes.c_uint64, ctypes.c_void_p]),
    "picha_b200_contribs": (ctypes.c_int, [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, _ip, _ip, _ip,
                                           _fp, _ip, ctypes.c_int]),
    "picha_b200_wide_blocks": (ctypes.c_int, [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, _fp, ctypes.c_int]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. picha_b200 has no CPU fallback: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C picha_b200/csrc).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


class PichaError(Exception):
    """A non-zero status from the C-ABI; ``str(e)`` is the reference's message where it has one."""

    def __init__(self, status):
        self.status = status
        msg = lib.picha_b200_strerror(status).decode()
        if status in (ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM):
            detail = lib.picha_b200_last_error().decode()
            if detail:
                msg = f"{msg}: {detail}"
        super().__init__(msg)


def check(status):
    if status != OK:
        raise PichaError(status)
    return status
