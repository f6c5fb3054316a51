"""CPU checkers for the picha pixel hot path.  TEST INFRASTRUCTURE ONLY.

Two interchangeable back ends behind one numpy-facing interface:

* ``port``  -- ``libpicha_oracle.so``: the plain-C restatement in ``picha_oracle.c``
  (citations to the reference's file:line live there);
* ``ref``   -- ``_ref/libpicha_ref.so``: the reference's own ``src/resize.cc`` and
  ``src/colorconvert.cc`` compiled from the read-only checkout by ``oracle/Makefile``
  (exists only where that build ran; it travels to the GPU box as a built file).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
import this package -- as the checker, never as the thing shipped or measured as the
product.  Nothing under ``picha_b200/`` imports it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PIXELS = ["rgb", "rgba", "grey", "greya", "r16", "r16g16", "r16g16b16", "r16g16b16a16"]
FILTERS = ["cubic", "lanczos", "catmulrom", "mitchel", "box", "triangle"]
PIXEL_BYTES = [3, 4, 1, 2, 2, 4, 6, 8]
PIXEL_CHANNELS = [3, 4, 1, 2, 1, 2, 3, 4]

_u8p = ctypes.POINTER(ctypes.c_ubyte)
_ip = ctypes.POINTER(ctypes.c_int)
_fp = ctypes.POINTER(ctypes.c_float)


def build(ref_root: str = "/root/reference") -> None:
    """Compile the checkers (``make -C oracle``).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", _HERE, f"REF={ref_root}"], check=True)


def _load(path):
    return ctypes.CDLL(path) if os.path.exists(path) else None


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(_HERE, "libpicha_oracle.so")
        if not os.path.exists(path):
            build()
        _port = ctypes.CDLL(path)
        _port.po_resize.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int]
        _port.po_color_convert.argtypes = [ctypes.c_float] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                                  ctypes.c_int, ctypes.c_int]
        _port.po_make_contribs.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                           _ip, _ip, _ip, _fp, ctypes.c_int]
        _port.po_resolve_color_settings.argtypes = [ctypes.c_double] * 3 + [_fp]
        _port.po_resolve_color_settings.restype = None
    return _port


def ref_lib():
    """The compiled reference, or None when it was never built here."""
    global _ref
    if _ref is None:
        lib = _load(os.path.join(_HERE, "_ref", "libpicha_ref.so"))
        if lib is None:
            return None
        lib.ref_resize.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int]
        lib.ref_resize.restype = None
        lib.ref_colorconvert.argtypes = [ctypes.c_float] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                                ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                                ctypes.c_int, ctypes.c_int]
        lib.ref_colorconvert.restype = None
        lib.ref_contribs.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                     _ip, _ip, _ip, _fp, ctypes.c_int]
        _ref = lib
    return _ref


def have_ref() -> bool:
    return ref_lib() is not None


def _impl(impl):
    """"auto" = the compiled reference (oracle/_ref) when it is present, else the C port."""
    if impl == "auto":
        return "ref" if have_ref() else "port"
    return impl


def _idx(name_or_idx, table):
    return table.index(name_or_idx) if isinstance(name_or_idx, str) else int(name_or_idx)


def row_stride(width, pixel):
    return (PIXEL_BYTES[_idx(pixel, PIXELS)] * width + 3) & ~3


def _check_buf(buf, stride, width, height, bpp):
    assert buf.dtype == np.uint8 and buf.flags["C_CONTIGUOUS"]
    assert stride >= width * bpp
    assert buf.size >= stride * (height - 1) + width * bpp


def resize(src, sstride, sw, sh, pixel, dw, dh, filt="cubic", fwidth=0.70, impl="auto", dstride=None,
           dst=None):
    """resizeImage on a flat uint8 buffer; returns (dst_buffer, dstride)."""
    impl = _impl(impl)
    p = _idx(pixel, PIXELS)
    f = _idx(filt, FILTERS)
    bpp = PIXEL_BYTES[p]
    _check_buf(src, sstride, sw, sh, bpp)
    if dstride is None:
        dstride = row_stride(dw, p)
    if dst is None:
        dst = np.zeros(dstride * dh, dtype=np.uint8)
    if impl == "ref":
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref/libpicha_ref.so not built")
        lib.ref_resize(f, fwidth, src.ctypes.data, sstride, sw, sh, dst.ctypes.data, dstride, dw, dh, p)
    else:
        rc = port_lib().po_resize(f, fwidth, src.ctypes.data, sstride, sw, sh, dst.ctypes.data, dstride, dw, dh, p)
        if rc != 0:
            raise ValueError(f"po_resize failed: {rc}")
    return dst, dstride


def resolve_color_settings(r=float("nan"), g=float("nan"), b=float("nan")):
    out = (ctypes.c_float * 3)()
    port_lib().po_resolve_color_settings(r, g, b, out)
    return float(out[0]), float(out[1]), float(out[2])


def color_convert(src, sstride, w, h, spixel, dpixel, weights=None, impl="auto", dstride=None, dst=None):
    """doColorConvert on a flat uint8 buffer; returns (dst_buffer, dstride)."""
    impl = _impl(impl)
    sp = _idx(spixel, PIXELS)
    dp = _idx(dpixel, PIXELS)
    _check_buf(src, sstride, w, h, PIXEL_BYTES[sp])
    if weights is None:
        weights = resolve_color_settings()
    if dstride is None:
        dstride = row_stride(w, dp)
    if dst is None:
        dst = np.zeros(dstride * h, dtype=np.uint8)
    r, g, b = weights
    if impl == "ref":
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref/libpicha_ref.so not built")
        lib.ref_colorconvert(r, g, b, src.ctypes.data, sstride, w, h, sp, dst.ctypes.data, dstride, dp)
    else:
        rc = port_lib().po_color_convert(r, g, b, src.ctypes.data, sstride, w, h, sp, dst.ctypes.data, dstride, dp)
        if rc != 0:
            raise ValueError(f"po_color_convert failed: {rc}")
    return dst, dstride


def cmyk_to_rgb(src, sstride, w, h, dstride=None):
    """cmyk_to_rgb of src/jpegcodec.cc:36-42 on a flat uint8 buffer of 4-byte pixels; returns (rgb_buffer, dstride)."""
    _check_buf(src, sstride, w, h, 4)
    if dstride is None:
        dstride = row_stride(w, 0)
    dst = np.zeros(dstride * h, dtype=np.uint8)
    lib = port_lib()
    lib.po_cmyk_to_rgb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.po_cmyk_to_rgb.restype = ctypes.c_int
    rc = lib.po_cmyk_to_rgb(src.ctypes.data, sstride, w, h, dst.ctypes.data, dstride)
    if rc != 0:
        raise ValueError(f"po_cmyk_to_rgb failed: {rc}")
    return dst, dstride


def contribs(filt, fwidth, srcsize, dstsize, impl="auto"):
    """One axis of makeContribs: (left[], right[], woff[], weights[])."""
    impl = _impl(impl)
    f = _idx(filt, FILTERS)
    left = np.zeros(dstsize, np.int32)
    right = np.zeros(dstsize, np.int32)
    woff = np.zeros(dstsize, np.int32)
    fn = ref_lib().ref_contribs if impl == "ref" else port_lib().po_make_contribs
    args = (f, fwidth, srcsize, dstsize, left.ctypes.data_as(_ip), right.ctypes.data_as(_ip),
            woff.ctypes.data_as(_ip))
    n = fn(*args, None, 0)
    w = np.zeros(max(n, 1), np.float32)
    fn(*args, w.ctypes.data_as(_fp), n)
    return left, right, woff, w[:n]


def payload(buf, stride, width, height, pixel):
    """Row payloads as an (h, w*bytes) uint8 view (what Image.row() exposes; padding excluded)."""
    bpp = PIXEL_BYTES[_idx(pixel, PIXELS)]
    rows = np.lib.stride_tricks.as_strided(buf, shape=(height, width * bpp), strides=(stride, 1))
    return rows


def channels(buf, stride, width, height, pixel):
    """Row payloads as channel values: uint8 or (for r16*) uint16, shape (h, w*channels)."""
    p = _idx(pixel, PIXELS)
    rows = np.ascontiguousarray(payload(buf, stride, width, height, p))
    return rows.view(np.uint16) if p >= 4 else rows
