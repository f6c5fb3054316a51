"""Regenerate the committed golden vectors (run in the build container only).

    python tests/golden/make_golden.py

Needs the read-only reference checkout at /root/reference and oracle/_ref/libpicha_ref.so
(``make -C oracle``).  Neither exists on the GPU box, so what this script writes is committed:

* ``picha_fixtures.npz`` -- the reference's own test fixtures decoded to raw pixel rows with PIL
  (test/resize.js:17-30 -> test2.jpg / test2.png; test/color_convert.js:16-28 -> test.png /
  greytest.png; test/codec.js:23-25 -> test.jpeg, the README resize example's input).
* ``ref_vectors.npz`` -- inputs and outputs of the reference's own C++ (resizeImage,
  doColorConvert, makeContribs) on a sweep the reference's tests never touch: every filter x
  every pixel format, up/down/integer/non-integer ratios, filterScale, strided inputs, custom
  luma weights, and the ring-aliasing cases (box at integer ratios).
"""
import itertools
import os
import sys

import numpy as np
from PIL import Image as PILImage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

REF_TEST = "/root/reference/test"


def decode(name, mode):
    im = PILImage.open(os.path.join(REF_TEST, name))
    assert im.mode == mode, (name, im.mode)
    return np.ascontiguousarray(np.asarray(im))


def fixtures():
    out = {
        "test2_jpg_rgb": decode("test2.jpg", "RGB"),        # 50 x 76 x 3
        "test2_png_rgb": decode("test2.png", "RGB"),        # 24 x 32 x 3
        "test_png_rgba": decode("test.png", "RGBA"),        # 50 x 50 x 4
        "greytest_png_greya": decode("greytest.png", "LA"),  # 50 x 50 x 2
        "test_jpeg_rgb": decode("test.jpeg", "RGB"),        # 50 x 50 x 3
    }
    np.savez_compressed(os.path.join(HERE, "picha_fixtures.npz"), **out)
    return out


RESIZE_SHAPES = [
    # sw, sh, dw, dh, filterScale, extra stride bytes
    (37, 29, 11, 7, 1.0, 0),
    (16, 12, 48, 40, 0.7, 4),
    (64, 48, 16, 12, 1.0, 0),     # integer 4x: lanczos 17 taps vs ring of 16
    (30, 30, 10, 10, 1.0, 8),     # integer 3x: box aliasing case
    (40, 20, 20, 10, 1.0, 0),     # integer 2x: box aliasing case
    (23, 17, 23, 17, 1.0, 0),     # identity size
    (9, 120, 9, 3, 1.0, 0),       # 40x vertical
    (25, 25, 50, 50, 0.7, 0),     # cfg1 in small (2x up, default cubic@0.70)
    (12, 12, 5, 7, 2.5, 0),
    (33, 21, 47, 9, 1.5, 4),      # up in x, down in y
]


def vectors(fx):
    rng = np.random.default_rng(20261018)
    out = {}
    meta = []
    k = 0
    for p, f in itertools.product(range(8), range(6)):
        for (sw, sh, dw, dh, fw, pad) in RESIZE_SHAPES:
            if (p + f + k) % 3 != 0 and not (f == 4 and fw == 1.0):
                k += 1
                continue          # thin the product: every (p, f) pair still gets >= 3 shapes
            ss = O.row_stride(sw, p) + pad
            src = rng.integers(0, 256, ss * sh, dtype=np.uint8)
            dst, ds = O.resize(src, ss, sw, sh, p, dw, dh, f, fw, impl="ref")
            out[f"rs{k}_src"] = src
            out[f"rs{k}_dst"] = O.payload(dst, ds, dw, dh, p).copy()
            meta.append((0, k, p, f, sw, sh, dw, dh, fw, ss))
            k += 1
    # colour conversion: all 64 pairs, default and custom weights, odd width, padded stride
    weights = [O.resolve_color_settings(), O.resolve_color_settings(0.2, 0.5, 0.3),
               O.resolve_color_settings(1, 1, 1)]
    k = 0
    for sp, dp in itertools.product(range(8), range(8)):
        for wi, wts in enumerate(weights):
            if wi > 0 and not (O.PIXEL_CHANNELS[sp] >= 3 and O.PIXEL_CHANNELS[dp] <= 2):
                continue          # weights only matter for luma pairs
            w, h = 53, 5
            ss = O.row_stride(w, sp) + 4
            src = rng.integers(0, 256, ss * h, dtype=np.uint8)
            dst, ds = O.color_convert(src, ss, w, h, sp, dp, wts, impl="ref")
            out[f"cc{k}_src"] = src
            out[f"cc{k}_dst"] = O.payload(dst, ds, w, h, dp).copy()
            meta.append((1, k, sp, dp, w, h, wi, 0, 0.0, ss))
            k += 1
    # exhaustive value tables through the reference: every u8 / u16 value, depth changes and luma extremes
    v8 = np.arange(256, dtype=np.uint8)
    v16 = np.arange(65536, dtype=np.uint16)
    d, _ = O.color_convert(v8.copy(), 256, 256, 1, "grey", "r16", impl="ref")
    out["tab_u8_to_u16"] = d[:512].view(np.uint16).copy()
    d, _ = O.color_convert(v16.view(np.uint8).copy(), 131072, 65536, 1, "r16", "grey", impl="ref")
    out["tab_u16_to_u8"] = d[:65536].copy()
    d, _ = O.color_convert(v16.view(np.uint8).copy(), 131072, 65536, 1, "r16", "r16g16", impl="ref")
    out["tab_u16_ident_fill"] = d[:262144].view(np.uint16).copy()
    # contribution tables for the benchmark shapes (left, right, weights)
    for name, (f, fw, s, dn) in {"cfg3x": (1, 1.0, 3840, 960), "cfg3y": (1, 1.0, 2160, 540),
                                 "cfg5x": (0, 0.7, 1920, 256), "cfg5y": (0, 0.7, 1080, 256),
                                 "cfg4": (3, 1.0, 2048, 4096), "cfg1": (0, 0.7, 50, 100)}.items():
        l, r, o, w = O.contribs(f, fw, s, dn, impl="ref")
        out[f"tab_{name}_left"], out[f"tab_{name}_right"], out[f"tab_{name}_w"] = l, r, w
    # the fixtures through the reference itself
    t2 = fx["test2_jpg_rgb"]
    h, w, _ = t2.shape
    ss = O.row_stride(w, "rgb")
    buf = np.zeros(ss * h, np.uint8)
    O.payload(buf, ss, w, h, "rgb")[:] = t2.reshape(h, -1)
    dst, ds = O.resize(buf, ss, w, h, "rgb", 32, 24, impl="ref")
    out["fixture_resize_ref"] = O.payload(dst, ds, 32, 24, "rgb").copy()
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    return len(meta)


if __name__ == "__main__":
    if not O.have_ref():
        raise SystemExit("oracle/_ref/libpicha_ref.so missing: run `make -C oracle` where /root/reference exists")
    fx = fixtures()
    n = vectors(fx)
    for f in ("picha_fixtures.npz", "ref_vectors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print("cases:", n)
