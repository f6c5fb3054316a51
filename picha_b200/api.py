"""Host-side mirror of picha's public resize / colorConvert surface (index.js:13-33) over the
C-ABI in include/picha_b200.h -- the same calls, option keys, defaults and error messages, so a
picha user (and picha's own tests, transliterated) finds the path unchanged:

    resize(image, opts, cb)            index.js:13   -> src/resize.cc:321  NAN_METHOD(resize)
    resizeSync(image, opts)            index.js:19   -> src/resize.cc:367  NAN_METHOD(resizeSync)
    colorConvert(image, opts, cb)      index.js:25   -> src/colorconvert.cc:215
    colorConvertSync(image, opts)      index.js:31   -> src/colorconvert.cc:257

Node is not available in the build image, so this layer is Python; the C++ addon glue that
does the same from JavaScript is in addon/ and INTEGRATION.md.  All pixel work happens in the
CUDA library; nothing here touches pixels.
"""
from __future__ import annotations

import ctypes
import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _native as N
from .image import Image, PIXEL_ENUM, PIXEL_NAMES

_pool = None
_EMPTY = np.zeros(16, dtype=np.uint8)   # stands in for the data pointer of zero-byte images


def _workers():
    """Async calls run on a small pool, as picha's run on libuv's (default 4 threads)."""
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix="picha_b200")
    return _pool


def _is_object(v):
    return isinstance(v, (dict, Image)) or hasattr(v, "__dict__")


def _get(obj, key):
    if isinstance(obj, dict):
        return obj.get(key)
    return getattr(obj, key, None)


def _to_uint32_as_int(v):
    """JS ToUint32 followed by the reference's assignment to `int` (src/resize.cc:341-342)."""
    if isinstance(v, bool) or not isinstance(v, (int, float)) or (isinstance(v, float) and not math.isfinite(v)):
        return 0
    u = int(v) & 0xFFFFFFFF
    return u - (1 << 32) if u >= (1 << 31) else u


def _number(v):
    """JS NumberValue: undefined/None and non-numbers become NaN."""
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        return float("nan")
    return float(v)


def _native_image(img):
    """jsImageToNativeImage (src/picha.cc:61-85): returns a CImage or None for "invalid image"."""
    pixel = PIXEL_ENUM.get(_get(img, "pixel"))
    data = _get(img, "data")
    if pixel is None or data is None:
        return None, None
    if not isinstance(data, np.ndarray):
        data = np.frombuffer(data, dtype=np.uint8)
    width = _to_uint32_as_int(_get(img, "width"))
    height = _to_uint32_as_int(_get(img, "height"))
    stride = _to_uint32_as_int(_get(img, "stride"))
    rw = N.lib.picha_b200_pixel_bytes(pixel) * width
    if height == 0 or width < 0 or height < 0 or data.nbytes < height * stride - stride + rw:   # src/picha.cc:78
        return None, None
    c = N.CImage(data.ctypes.data, stride, width, height, pixel)
    return c, data   # keep `data` alive for the duration of the call


def _new_image(width, height, pixel):
    """newJsImage (src/picha.cc:119-133): row_stride, buffer of height*stride bytes."""
    stride = N.lib.picha_b200_row_stride(width, pixel)
    data = np.zeros(stride * height, dtype=np.uint8)
    img = Image({"width": width, "height": height, "pixel": PIXEL_NAMES[pixel], "stride": stride, "data": data})
    base = data.ctypes.data if data.size else _EMPTY.ctypes.data   # never NULL; rows of zero bytes touch nothing
    return img, N.CImage(base, stride, width, height, pixel)


def _resize_options(opts):
    """getResizeOptions (src/resize.cc:179-198)."""
    filt = _get(opts, "filter")
    scale = _get(opts, "filterScale")
    has_filter = filt is not None
    tag = -1
    if has_filter:
        tag = N.FILTERS.index(filt) if (isinstance(filt, str) and filt in N.FILTERS) else -1
    tag_out = ctypes.c_int(0)
    width_out = ctypes.c_float(0)
    rc = N.lib.picha_b200_resolve_resize_options(int(has_filter), tag, int(scale is not None), _number(scale),
                                                 ctypes.byref(tag_out), ctypes.byref(width_out))
    if rc:
        raise N.PichaError(rc)
    return tag_out.value, width_out.value


def _prepare_resize(img, opts):
    src, keep = _native_image(img)
    if src is None:
        raise N.PichaError(N.ERR_INVALID_IMAGE)
    width = _to_uint32_as_int(_get(opts, "width"))
    height = _to_uint32_as_int(_get(opts, "height"))
    if width <= 0 or height <= 0:
        raise N.PichaError(N.ERR_INVALID_DIMENSIONS)
    tag, fwidth = _resize_options(opts)
    out, dst = _new_image(width, height, src.pixel)
    # extensions to picha's option object: exact=True forces the bit-exact kernel, fast=True the
    # throughput kernel (by default small images get the former, large ones the latter)
    flags = N.EXACT if _get(opts, "exact") else (N.FORCE_FAST if _get(opts, "fast") else 0)
    return src, keep, out, dst, tag, fwidth, flags


def resizeSync(img, opts):
    if not _is_object(img) or not _is_object(opts):
        raise TypeError("expected: resizeSync(image, opts)")
    src, keep, out, dst, tag, fwidth, flags = _prepare_resize(img, opts)
    N.check(N.lib.picha_b200_resize_ex(ctypes.byref(src), ctypes.byref(dst), tag, fwidth, flags))
    return out


def resize(img, opts, cb):
    if not _is_object(img) or not _is_object(opts) or not callable(cb):
        raise TypeError("expected: resize(image, opts, cb)")
    src, keep, out, dst, tag, fwidth, flags = _prepare_resize(img, opts)

    def work():
        try:
            N.check(N.lib.picha_b200_resize_ex(ctypes.byref(src), ctypes.byref(dst), tag, fwidth, flags))
        except Exception as e:   # the reference's core cannot fail; CUDA can
            cb(e, None)
            return
        _ = keep
        cb(None, out)

    return _workers().submit(work)


def _color_settings(opts):
    out = (ctypes.c_float * 3)()
    N.lib.picha_b200_resolve_color_settings(_number(_get(opts, "redWeight")), _number(_get(opts, "greenWeight")),
                                            _number(_get(opts, "blueWeight")), out)
    return out[0], out[1], out[2]


def _prepare_convert(img, opts):
    src, keep = _native_image(img)
    if src is None:
        raise N.PichaError(N.ERR_INVALID_IMAGE)
    to = PIXEL_ENUM.get(_get(opts, "pixel")) if isinstance(_get(opts, "pixel"), str) else None
    if to is None:
        raise N.PichaError(N.ERR_INVALID_PIXEL)
    out, dst = _new_image(src.width, src.height, to)
    return src, keep, out, dst, _color_settings(opts)


def colorConvertSync(img, opts):
    if not _is_object(img) or not _is_object(opts):
        raise TypeError("expected: colorConvertSync(image, opts)")
    src, keep, out, dst, (r, g, b) = _prepare_convert(img, opts)
    N.check(N.lib.picha_b200_color_convert(ctypes.byref(src), ctypes.byref(dst), r, g, b))
    return out


def resizeConvertSync(img, opts):
    """colorConvertSync(resizeSync(img, opts), opts) in one kernel (include/picha_b200.h: picha_b200_resize_convert):
    `opts` carries resize's keys (width, height, filter, filterScale) and colorConvert's (pixel, redWeight, ...).
    Not part of picha's JS surface; the fused form of the pair a thumbnailer calls back to back."""
    if not _is_object(img) or not _is_object(opts):
        raise TypeError("expected: resizeConvertSync(image, opts)")
    src, keep, _, _, tag, fwidth, flags = _prepare_resize(img, opts)
    to = PIXEL_ENUM.get(_get(opts, "pixel")) if isinstance(_get(opts, "pixel"), str) else None
    if to is None:
        raise N.PichaError(N.ERR_INVALID_PIXEL)
    out, dst = _new_image(_to_uint32_as_int(_get(opts, "width")), _to_uint32_as_int(_get(opts, "height")), to)
    r, g, b = _color_settings(opts)
    N.check(N.lib.picha_b200_resize_convert(ctypes.byref(src), ctypes.byref(dst), tag, fwidth, r, g, b, flags))
    return out


def cmykToRgbSync(img):
    """The JPEG decoder's cmyk_to_rgb (src/jpegcodec.cc:36-42) on a whole image: `img` carries C, M, Y, K in
    an 'rgba' image; the result is 'rgb'.  Not part of picha's JS surface (the reference runs this loop
    inside decodeJpeg); exposed for the step in front of the hot path."""
    if not _is_object(img):
        raise TypeError("expected: cmykToRgbSync(image)")
    src, keep, out, dst, _ = _prepare_convert(img, {"pixel": "rgb"})
    if src.pixel != N.PIXELS.index("rgba"):
        raise N.PichaError(N.ERR_FORMAT_MISMATCH)
    N.check(N.lib.picha_b200_cmyk_to_rgb(ctypes.byref(src), ctypes.byref(dst)))
    _ = keep
    return out


def colorConvert(img, opts, cb):
    if not _is_object(img) or not _is_object(opts) or not callable(cb):
        raise TypeError("expected: colorConvert(image, opts, cb)")
    src, keep, out, dst, (r, g, b) = _prepare_convert(img, opts)

    def work():
        try:
            N.check(N.lib.picha_b200_color_convert(ctypes.byref(src), ctypes.byref(dst), r, g, b))
        except Exception as e:
            cb(e, None)
            return
        _ = keep
        cb(None, out)

    return _workers().submit(work)


# ---- data-parallel batches (no counterpart in index.js: a JS caller gets the same effect by
#      issuing many picha.resize calls; the addon can forward them here) -----------------------

def _c_array(items):
    arr = (N.CImage * len(items))()
    for i, c in enumerate(items):
        arr[i] = c
    return arr


def resizeBatchSync(images, opts, device=-1):
    """Resize independent images; device=-1 shards them over every GPU of the box."""
    preps = [_prepare_resize(im, opts) for im in images]
    if not preps:
        return []
    srcs = _c_array([p[0] for p in preps])
    dsts = _c_array([p[3] for p in preps])
    _, _, _, _, tag, fwidth, flags = preps[0]
    N.check(N.lib.picha_b200_resize_batch(len(preps), srcs, dsts, tag, fwidth, flags, device))
    return [p[2] for p in preps]


def colorConvertBatchSync(images, opts, device=-1):
    preps = [_prepare_convert(im, opts) for im in images]
    if not preps:
        return []
    srcs = _c_array([p[0] for p in preps])
    dsts = _c_array([p[3] for p in preps])
    r, g, b = preps[0][4]
    N.check(N.lib.picha_b200_color_convert_batch(len(preps), srcs, dsts, r, g, b, device))
    return [p[2] for p in preps]
