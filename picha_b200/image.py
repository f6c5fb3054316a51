"""``Image`` -- Python mirror of picha's JS Image class (lib/image.js), the object every
picha call takes and returns.  Pure host logic; semantics follow the reference line by line
so the parity tests read like test/*.js:

* ctor validation and default stride ``(width*psize + 3) & ~3``     lib/image.js:3-21
* ``pixelSize`` table -- including the reference's ``'r16b16'`` spelling of the 4-byte
  16-bit format; ``'r16g16'`` (the native name, src/picha.h:153) is accepted too
* ``row``, ``equalPixels``, ``avgChannelDiff`` compare row payloads, never padding   :42-74
* ``subView`` shares the buffer and keeps the parent's stride        :76-87
* ``copy`` copies the overlapping payload                            :89-96
"""
from __future__ import annotations

import numpy as np

PIXEL_SIZES = {
    "rgb": 3, "rgba": 4, "grey": 1, "greya": 2,
    "r16g16b16": 6, "r16g16b16a16": 8, "r16": 2, "r16b16": 4,
    "r16g16": 4,
}
# name -> enum PixelMode (src/picha.h:79-92)
PIXEL_ENUM = {"rgb": 0, "rgba": 1, "grey": 2, "greya": 3, "r16": 4, "r16g16": 5, "r16b16": 5,
              "r16g16b16": 6, "r16g16b16a16": 7}
PIXEL_NAMES = ["rgb", "rgba", "grey", "greya", "r16", "r16g16", "r16g16b16", "r16g16b16a16"]


class Image:
    def __init__(self, opt=None, **kw):
        opt = dict(opt or {}, **kw)
        self.data = opt.get("data")
        self.width = opt.get("width") or 0
        self.height = opt.get("height") or 0
        self.pixel = opt.get("pixel") or "rgba"
        psize = Image.pixelSize(self.pixel)
        self.stride = opt.get("stride") or ((self.width * psize + 3) & ~3)
        if psize == 0:
            raise ValueError("invalid pixel format " + str(self.pixel))
        if self.stride < self.width * psize:
            raise ValueError("stride too short")
        if self.width < 0 or self.height < 0:
            raise ValueError("invalid dimensions")
        if self.stride * self.height != 0 and self.data is None:
            self.data = np.zeros(self.stride * self.height, dtype=np.uint8)   # Buffer.alloc zero-fills
        if self.data is not None:
            if not isinstance(self.data, np.ndarray):
                self.data = np.frombuffer(self.data, dtype=np.uint8)
            if self.data.dtype != np.uint8 or self.data.ndim != 1:
                self.data = self.data.reshape(-1).view(np.uint8)
            if self.data.size < self.stride * (self.height - 1) + self.width * psize:
                raise ValueError("image data too small")

    @staticmethod
    def pixelSize(pixel=None):
        return PIXEL_SIZES.get(pixel, 0)

    def pixel_size(self):
        return PIXEL_SIZES.get(self.pixel, 0)

    def row(self, y):
        o = y * self.stride
        return self.data[o:o + self.width * self.pixel_size()]

    def rows(self):
        """(height, width*bytes) strided view of the payload."""
        return np.lib.stride_tricks.as_strided(self.data, shape=(self.height, self.width * self.pixel_size()),
                                               strides=(self.stride, 1), writeable=False)

    def equalPixels(self, o):
        if self.width != o.width or self.height != o.height or self.pixel != o.pixel:
            return False
        return all(np.array_equal(self.row(y), o.row(y)) for y in range(self.height))

    def avgChannelDiff(self, o):
        if self.width != o.width or self.height != o.height or self.pixel != o.pixel:
            return 255
        rw = self.width * self.pixel_size()
        s = np.abs(self.rows().astype(np.int64) - o.rows().astype(np.int64)).sum()
        return float(s) / (self.height * rw)

    def channelDiff16(self, o):
        """(max, mean) absolute difference on channel values -- uint16 for the r16* formats, where
        a one-step difference across a byte boundary would read as 255 in avgChannelDiff."""
        a, b = self.rows(), o.rows()
        if PIXEL_ENUM[self.pixel] >= 4:
            a = np.ascontiguousarray(a).view(np.uint16)
            b = np.ascontiguousarray(b).view(np.uint16)
        d = np.abs(a.astype(np.int64) - b.astype(np.int64))
        return int(d.max()) if d.size else 0, float(d.mean()) if d.size else 0.0

    def subView(self, x, y, w, h):
        p = self.pixel_size()
        off = y * self.stride + x * p
        length = (h - 1) * self.stride + w * p
        return Image({"width": w, "height": h, "pixel": self.pixel, "stride": self.stride,
                      "data": self.data[off:off + length]})

    def copy(self, target):
        if target.pixel != self.pixel:
            raise ValueError("can't copy pixels between different pixel types")
        rw = self.pixel_size() * min(self.width, target.width)
        h = min(self.height, target.height)
        for y in range(h):
            target.data[y * target.stride:y * target.stride + rw] = self.data[y * self.stride:y * self.stride + rw]

    # camelCase aliases keep call sites identical to the JS tests
    pixelSizeOf = pixel_size
