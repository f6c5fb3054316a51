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
