"""Device-resident batches: the data-parallel form of the hot path.

A ``DeviceBatch`` is n same-shape images in one torch uint8 CUDA tensor (torch is only the
allocator and the stream here); ``resize`` / ``color_convert`` call the C-ABI's *_device entry
points on torch's current stream, so nothing crosses PCIe.  Images are independent: ranks of
a multi-GPU job each own a block of the batch and never exchange anything.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native as N
from .image import Image, PIXEL_ENUM, PIXEL_NAMES, PIXEL_SIZES


def _align(v, a):
    return (v + a - 1) // a * a


class DeviceBatch:
    def __init__(self, n, width, height, pixel, stride=None, device=None, pitch_align=128):
        self.n, self.width, self.height, self.pixel = n, width, height, pixel
        bpp = PIXEL_SIZES[pixel]
        self.stride = stride or _align(width * bpp, pitch_align)
        self.step = _align(self.stride * height, 256)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.buf = torch.empty(max(n * self.step, 1), dtype=torch.uint8, device=self.device)

    def cimage(self, index=0):
        return N.CImage(self.buf.data_ptr() + index * self.step, self.stride, self.width, self.height,
                        PIXEL_ENUM[self.pixel])

    @property
    def payload_bytes(self):
        return self.n * self.width * self.height * PIXEL_SIZES[self.pixel]

    def fill_synthetic(self, seed, first_image=0):
        c = self.cimage()
        with torch.cuda.device(self.device):
            N.check(N.lib.picha_b200_synthetic_fill_device(self.n, ctypes.byref(c), self.step, seed, first_image,
                                                           torch.cuda.current_stream().cuda_stream))

    def image(self, index):
        """Copy image `index` back to the host as an Image (parity checks)."""
        raw = self.buf[index * self.step:index * self.step + self.stride * self.height].cpu().numpy()
        return Image({"width": self.width, "height": self.height, "pixel": self.pixel, "stride": self.stride,
                      "data": raw})

    def upload(self, index, image):
        assert image.width == self.width and image.height == self.height
        rows = torch.from_numpy(image.rows().copy())
        view = self.buf[index * self.step:index * self.step + self.stride * self.height].view(self.height, self.stride)
        view[:, :rows.shape[1]] = rows.to(self.device)


def resize(src, dst, filter=None, filter_scale=None, exact=False, fast=False):
    """dst[i] = resizeImage(src[i]) for every image of the batch, on the current stream.  Options resolve like
    getResizeOptions (src/resize.cc:173-198): no filter key = cubic at width 0.70, a named filter = width 1.0."""
    tag_out, width_out = ctypes.c_int(0), ctypes.c_float(0)
    has_filter = filter is not None
    N.check(N.lib.picha_b200_resolve_resize_options(int(has_filter), N.FILTERS.index(filter) if has_filter else 0,
                                                    int(filter_scale is not None),
                                                    float(filter_scale if filter_scale is not None else 0.0),
                                                    ctypes.byref(tag_out), ctypes.byref(width_out)))
    s, d = src.cimage(), dst.cimage()
    with torch.cuda.device(src.device):
        N.check(N.lib.picha_b200_resize_device(src.n, ctypes.byref(s), src.step, ctypes.byref(d), dst.step,
                                               tag_out.value, width_out.value,
                                               N.EXACT if exact else (N.FORCE_FAST if fast else 0),
                                               torch.cuda.current_stream().cuda_stream))


def color_convert(src, dst, weights=None):
    out = (ctypes.c_float * 3)()
    nan = float("nan")
    r, g, b = weights if weights is not None else (nan, nan, nan)
    N.lib.picha_b200_resolve_color_settings(r, g, b, out)
    s, d = src.cimage(), dst.cimage()
    with torch.cuda.device(src.device):
        N.check(N.lib.picha_b200_color_convert_device(src.n, ctypes.byref(s), src.step, ctypes.byref(d), dst.step,
                                                      out[0], out[1], out[2],
                                                      torch.cuda.current_stream().cuda_stream))


def resize_convert(src, dst, filter=None, filter_scale=None, weights=None, exact=False):
    """dst[i] = doColorConvert(resizeImage(src[i])) in one kernel; dst has the target size AND pixel format."""
    tag_out, width_out = ctypes.c_int(0), ctypes.c_float(0)
    has_filter = filter is not None
    N.check(N.lib.picha_b200_resolve_resize_options(int(has_filter), N.FILTERS.index(filter) if has_filter else 0,
                                                    int(filter_scale is not None),
                                                    float(filter_scale if filter_scale is not None else 0.0),
                                                    ctypes.byref(tag_out), ctypes.byref(width_out)))
    out = (ctypes.c_float * 3)()
    nan = float("nan")
    r, g, b = weights if weights is not None else (nan, nan, nan)
    N.lib.picha_b200_resolve_color_settings(r, g, b, out)
    s, d = src.cimage(), dst.cimage()
    with torch.cuda.device(src.device):
        N.check(N.lib.picha_b200_resize_convert_device(src.n, ctypes.byref(s), src.step, ctypes.byref(d), dst.step,
                                                       tag_out.value, width_out.value, out[0], out[1], out[2],
                                                       N.EXACT if exact else 0, torch.cuda.current_stream().cuda_stream))


class DeviceImage:
    """picha's Image (lib/image.js) with its pixels resident in HBM: the same fields (data, width, height, pixel,
    stride) and the same stride semantics, `data` being a 1-D torch uint8 CUDA tensor instead of a Buffer.
    ``row`` / ``subView`` share the storage exactly like Buffer.slice does (lib/image.js:42-44,76-87), ``copy`` copies
    the overlapping payload (:89-96) with one strided device copy; ``resize`` / ``colorConvert`` / ``resizeConvert`` call
    the C-ABI's *_device entry points on torch's current stream -- a chain such as
    ``img.resize(...).subView(...).colorConvert(...)`` never crosses PCIe.
    """

    def __init__(self, opt=None, **kw):
        opt = dict(opt or {}, **kw)
        self.width, self.height = opt.get("width") or 0, opt.get("height") or 0
        self.pixel = opt.get("pixel") or "rgba"
        psize = PIXEL_SIZES.get(self.pixel, 0)
        if psize == 0:
            raise ValueError("invalid pixel format " + str(self.pixel))
        # default pitch: 128-byte rows (every row line-aligned and TMA-addressable); any stride >= width*bytes is accepted
        self.stride = opt.get("stride") or _align(self.width * psize, 128)
        if self.stride < self.width * psize:
            raise ValueError("stride too short")
        if self.width < 0 or self.height < 0:
            raise ValueError("invalid dimensions")
        self.data = opt.get("data")
        if self.data is None:
            dev = opt.get("device") or torch.device("cuda", torch.cuda.current_device())
            self.data = torch.empty(max(self.stride * self.height, 1), dtype=torch.uint8, device=dev)
        if self.data.dtype != torch.uint8 or self.data.dim() != 1 or not self.data.is_cuda:
            raise ValueError("data must be a 1-D uint8 CUDA tensor")
        if self.data.numel() < self.stride * (self.height - 1) + self.width * psize:
            raise ValueError("image data too small")

    # ---- lib/image.js semantics -----------------------------------------------------------------
    def pixelSize(self):
        return PIXEL_SIZES[self.pixel]

    def row(self, y):
        return self.data[y * self.stride:y * self.stride + self.width * self.pixelSize()]

    def rows(self):
        """(height, width*bytes) strided view of the payload (shares storage)."""
        return torch.as_strided(self.data, (self.height, self.width * self.pixelSize()), (self.stride, 1))

    def subView(self, x, y, w, h):
        p = self.pixelSize()
        off = y * self.stride + x * p
        return DeviceImage({"width": w, "height": h, "pixel": self.pixel, "stride": self.stride,
                            "data": self.data[off:off + (h - 1) * self.stride + w * p]})

    def copy(self, target):
        if target.pixel != self.pixel:
            raise ValueError("can't copy pixels between different pixel types")
        w, h = min(self.width, target.width), min(self.height, target.height)
        if w and h:
            rw = w * self.pixelSize()
            torch.as_strided(target.data, (h, rw), (target.stride, 1)).copy_(torch.as_strided(self.data, (h, rw), (self.stride, 1)))

    def equalPixels(self, o):
        if self.width != o.width or self.height != o.height or self.pixel != o.pixel:
            return False
        return bool(torch.equal(self.rows(), o.rows()))

    # ---- host <-> device ------------------------------------------------------------------------
    @staticmethod
    def from_host(image, device=None):
        out = DeviceImage({"width": image.width, "height": image.height, "pixel": image.pixel, "device": device})
        if image.width and image.height:
            out.rows().copy_(torch.from_numpy(image.rows().copy()))
        return out

    def to_host(self):
        out = Image({"width": self.width, "height": self.height, "pixel": self.pixel})
        if self.width and self.height:
            rw = self.width * self.pixelSize()
            import numpy as np
            np.lib.stride_tricks.as_strided(out.data, shape=(self.height, rw), strides=(out.stride, 1))[:] = self.rows().cpu().numpy()
        return out

    # ---- the hot path, on the device ----------------------------------------------------------------
    def _cimage(self):
        return N.CImage(self.data.data_ptr(), self.stride, self.width, self.height, PIXEL_ENUM[self.pixel])

    def _resize_options(self, opts):
        tag, width = ctypes.c_int(0), ctypes.c_float(0)
        f, fs = opts.get("filter"), opts.get("filterScale")
        if f is not None and f not in N.FILTERS:
            raise N.PichaError(N.ERR_INVALID_FILTER)
        N.check(N.lib.picha_b200_resolve_resize_options(int(f is not None), N.FILTERS.index(f) if f is not None else 0,
                                                        int(fs is not None), float(fs if fs is not None else 0.0),
                                                        ctypes.byref(tag), ctypes.byref(width)))
        return tag.value, width.value

    @staticmethod
    def _luma(opts):
        out = (ctypes.c_float * 3)()
        nan = float("nan")
        N.lib.picha_b200_resolve_color_settings(opts.get("redWeight", nan), opts.get("greenWeight", nan), opts.get("blueWeight", nan), out)
        return out[0], out[1], out[2]

    def resize(self, opts):
        """picha.resizeSync(image, opts) (index.js:19-21) without leaving the device."""
        tag, fwidth = self._resize_options(opts)
        dst = DeviceImage({"width": int(opts["width"]), "height": int(opts["height"]), "pixel": self.pixel, "device": self.data.device})
        s, d = self._cimage(), dst._cimage()
        with torch.cuda.device(self.data.device):
            N.check(N.lib.picha_b200_resize_device(1, ctypes.byref(s), 0, ctypes.byref(d), 0, tag, fwidth,
                                                   N.EXACT if opts.get("exact") else 0, torch.cuda.current_stream().cuda_stream))
        return dst

    def colorConvert(self, opts):
        """picha.colorConvertSync(image, opts) (index.js:31-33) without leaving the device."""
        if opts.get("pixel") not in PIXEL_ENUM:
            raise N.PichaError(N.ERR_INVALID_PIXEL)
        dst = DeviceImage({"width": self.width, "height": self.height, "pixel": opts["pixel"], "device": self.data.device})
        r, g, b = self._luma(opts)
        s, d = self._cimage(), dst._cimage()
        with torch.cuda.device(self.data.device):
            N.check(N.lib.picha_b200_color_convert_device(1, ctypes.byref(s), 0, ctypes.byref(d), 0, r, g, b,
                                                          torch.cuda.current_stream().cuda_stream))
        return dst

    def resizeConvert(self, opts):
        """colorConvert(resize(image)) in one kernel (picha_b200_resize_convert_device)."""
        if opts.get("pixel") not in PIXEL_ENUM:
            raise N.PichaError(N.ERR_INVALID_PIXEL)
        tag, fwidth = self._resize_options(opts)
        dst = DeviceImage({"width": int(opts["width"]), "height": int(opts["height"]), "pixel": opts["pixel"], "device": self.data.device})
        r, g, b = self._luma(opts)
        s, d = self._cimage(), dst._cimage()
        with torch.cuda.device(self.data.device):
            N.check(N.lib.picha_b200_resize_convert_device(1, ctypes.byref(s), 0, ctypes.byref(d), 0, tag, fwidth, r, g, b,
                                                           N.EXACT if opts.get("exact") else 0, torch.cuda.current_stream().cuda_stream))
        return dst
