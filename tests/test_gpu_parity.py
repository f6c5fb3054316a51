"""Parity of the CUDA path with the oracle, through the C-ABI (run with -m gpu on a B200).

colorConvert: bit-exact.  resize: bit-exact in EXACT mode; the default (fast) mode must be within
+-1 LSB per channel and mean |diff| <= 0.05 (uint16 steps for the r16* formats) -- the tolerance
BASELINE.json's north_star states.
"""
import ctypes
import itertools
import threading

import numpy as np
import pytest

import oracle as O
from picha_b200 import _native as N
from picha_b200.image import Image, PIXEL_NAMES
from picha_b200.synthetic import fill_host

pytestmark = pytest.mark.gpu

MAX_LSB = 1
MEAN_LSB = 0.05
# the bit-exact kernel, the default choice, and the throughput kernel forced onto small images
MODES = [{"exact": True}, {}, {"fast": True}]


def rand_image(rng, w, h, pixel, pad=0, offset=0):
    bpp = O.PIXEL_BYTES[O.PIXELS.index(pixel)]
    stride = ((w * bpp + 3) & ~3) + pad
    raw = rng.integers(0, 256, stride * h + offset, dtype=np.uint8)
    return Image({"width": w, "height": h, "pixel": pixel, "stride": stride, "data": raw[offset:]})


def chan(img):
    r = np.ascontiguousarray(img.rows())
    return r.view(np.uint16) if PIXEL_NAMES.index(img.pixel if img.pixel != "r16b16" else "r16g16") >= 4 else r


def oracle_resize(img, dw, dh, filt, fw):
    d, ds = O.resize(np.ascontiguousarray(img.data), img.stride, img.width, img.height, img.pixel, dw, dh, filt, fw)
    return Image({"width": dw, "height": dh, "pixel": img.pixel, "stride": ds, "data": d})


def assert_resize_close(got, want, exact, ctx):
    """exact: identical.  Otherwise max |diff| <= 1 step and mean |diff| <= 0.05 steps (u8 steps, or
    u16 steps for the r16* formats).  On outputs of fewer than 4096 values a single off-by-one would
    already break the mean, so there the mean bound reads "at most max(1, 5 %) of the values differ"."""
    a, b = chan(got).astype(np.int64), chan(want).astype(np.int64)
    d = np.abs(a - b)
    if exact:
        assert d.max() == 0, (ctx, "exact mode differs", int(d.max()), float(d.mean()))
        return
    assert d.max() <= MAX_LSB, (ctx, int(d.max()), float(d.mean()))
    if d.size >= 4096:
        assert d.mean() <= MEAN_LSB, (ctx, int(d.max()), float(d.mean()))
    else:
        assert int((d > 0).sum()) <= max(1, int(MEAN_LSB * d.size)), (ctx, int((d > 0).sum()), d.size)


# ---- the reference's own tests, transliterated -------------------------------------------------

def test_resize_fixture_sync_and_async(gpu, fixtures):
    """test/resize.js:17-30."""
    P = gpu
    rows = fixtures["test2_jpg_rgb"]
    h, w, _ = rows.shape
    image = Image({"width": w, "height": h, "pixel": "rgb"})
    for y in range(h):
        image.row(y)[:] = rows[y].reshape(-1)
    small = Image({"width": 32, "height": 24, "pixel": "rgb"})
    for y in range(24):
        small.row(y)[:] = fixtures["test2_png_rgb"][y].reshape(-1)
    opts = {"width": 32, "height": 24}
    box = {}
    done = threading.Event()

    def cb(err, o):
        box["err"], box["img"] = err, o
        done.set()

    P.resize(image, opts, cb)
    assert done.wait(60) and box["err"] is None
    async_small = box["img"]
    assert async_small.avgChannelDiff(small) < 2
    sync_small = P.resizeSync(image, opts)
    assert sync_small.avgChannelDiff(small) < 2
    assert sync_small.equalPixels(async_small)
    # stronger than the reference's own bound: north_star tolerance, and bit-exact in exact mode
    assert_resize_close(sync_small, small, False, "fixture")
    assert P.resizeSync(image, dict(opts, exact=True)).equalPixels(small)


def test_colour_fixture(gpu, fixtures):
    """test/color_convert.js:16-39."""
    P = gpu
    rgba = Image({"width": 50, "height": 50, "pixel": "rgba", "data": fixtures["test_png_rgba"].reshape(-1).copy()})
    grey = Image({"width": 50, "height": 50, "pixel": "greya", "data": fixtures["greytest_png_greya"].reshape(-1).copy()})
    to_grey = P.colorConvertSync(rgba, {"pixel": "greya"})
    assert to_grey.pixel == "greya" and to_grey.width == 50 and to_grey.height == 50
    assert to_grey.equalPixels(grey)
    box = {}
    done = threading.Event()

    def cb2(err, rimg):
        box["err"], box["img"] = err, rimg
        done.set()

    def cb1(err, img):
        assert err is None
        P.colorConvert(img, {"pixel": grey.pixel}, cb2)

    P.colorConvert(grey, {"pixel": "rgba"}, cb1)
    assert done.wait(60) and box["err"] is None
    assert grey.equalPixels(box["img"])


def test_readme_example_cfg1(gpu, fixtures):
    """BASELINE cfg1 / README.md:33-34: test.jpeg decoded to rgb (50x50), resizeSync to 100x100, default options."""
    P = gpu
    rows = fixtures["test_jpeg_rgb"]
    h, w, _ = rows.shape
    image = Image({"width": w, "height": h, "pixel": "rgb"})
    for y in range(h):
        image.row(y)[:] = rows[y].reshape(-1)
    want = oracle_resize(image, 100, 100, "cubic", np.float32(0.70))
    got = P.resizeSync(image, {"width": 100, "height": 100})
    assert got.width == 100 and got.height == 100 and got.pixel == "rgb" and got.stride == 300
    assert got.equalPixels(want)                       # small image: the default is the bit-exact kernel
    assert_resize_close(P.resizeSync(image, {"width": 100, "height": 100, "fast": True}), want, False, "cfg1 fast")


def test_committed_reference_vectors(gpu, ref_vectors):
    """Outputs of the reference's own C++ (tests/golden/ref_vectors.npz) straight against the GPU."""
    P = gpu
    weights = [None, (0.2, 0.5, 0.3), (1, 1, 1)]
    for row in ref_vectors["meta"]:
        kind, k = int(row[0]), int(row[1])
        if kind == 0:
            p, f, sw, sh, dw, dh = (int(v) for v in row[2:8])
            fw, ss = float(row[8]), int(row[9])
            img = Image({"width": sw, "height": sh, "pixel": PIXEL_NAMES[p], "stride": ss, "data": ref_vectors[f"rs{k}_src"].copy()})
            want = ref_vectors[f"rs{k}_dst"]
            ref = Image({"width": dw, "height": dh, "pixel": PIXEL_NAMES[p], "stride": want.shape[1], "data": want.reshape(-1).copy()})
            for mode in MODES:
                got = P.resizeSync(img, dict({"width": dw, "height": dh, "filter": N.FILTERS[f], "filterScale": fw}, **mode))
                assert_resize_close(got, ref, mode.get("exact", False), ("vec", k, p, f, sw, sh, dw, dh, fw, mode))
        else:
            sp, dp, w, h, wi = (int(v) for v in row[2:7])
            ss = int(row[9])
            img = Image({"width": w, "height": h, "pixel": PIXEL_NAMES[sp], "stride": ss, "data": ref_vectors[f"cc{k}_src"].copy()})
            opts = {"pixel": PIXEL_NAMES[dp]}
            if weights[wi]:
                opts.update(redWeight=weights[wi][0], greenWeight=weights[wi][1], blueWeight=weights[wi][2])
            got = P.colorConvertSync(img, opts)
            assert np.array_equal(got.rows(), ref_vectors[f"cc{k}_dst"]), ("vec cc", k, sp, dp, wi)


# ---- colour conversion: all 64 pairs, bit-exact ----------------------------------------------

@pytest.mark.parametrize("sp", range(8))
def test_color_convert_all_pairs_bit_exact(gpu, sp):
    P = gpu
    rng = np.random.default_rng(100 + sp)
    # 300 px: two full 128-pixel warp steps + a tail; 128: exactly one step; 5: tail only
    for dp, (w, h, pad, off) in itertools.product(range(8), [(300, 9, 0, 0), (128, 3, 8, 0), (5, 4, 0, 0), (131, 5, 0, 3)]):
        img = rand_image(rng, w, h, PIXEL_NAMES[sp], pad, off)
        want, ws = O.color_convert(np.ascontiguousarray(img.data), img.stride, w, h, sp, dp)
        got = P.colorConvertSync(img, {"pixel": PIXEL_NAMES[dp]})
        assert np.array_equal(got.rows(), O.payload(want, ws, w, h, dp)), (sp, dp, w, h, pad, off)


def test_color_convert_exhaustive_values(gpu):
    """Every u8 and u16 channel value through the depth-changing and luma paths."""
    P = gpu
    v16 = np.arange(65536, dtype=np.uint16)
    img = Image({"width": 65536, "height": 1, "pixel": "r16", "data": v16.view(np.uint8).copy()})
    for to in ("grey", "greya", "rgb", "r16g16b16a16", "rgba"):
        want, ws = O.color_convert(img.data, img.stride, 65536, 1, "r16", to)
        assert np.array_equal(P.colorConvertSync(img, {"pixel": to}).rows(), O.payload(want, ws, 65536, 1, to)), to
    v8 = np.arange(256, dtype=np.uint8)
    img = Image({"width": 256, "height": 1, "pixel": "grey", "data": v8.copy()})
    for to in ("r16", "r16g16", "r16g16b16", "r16g16b16a16", "rgba"):
        want, ws = O.color_convert(img.data, img.stride, 256, 1, "grey", to)
        assert np.array_equal(P.colorConvertSync(img, {"pixel": to}).rows(), O.payload(want, ws, 256, 1, to)), to
    # luma over a dense lattice of (r, g, b): FMA contraction would flip some of these (SURVEY item 4)
    r, g, b = np.meshgrid(np.arange(0, 256, 3), np.arange(0, 256, 5), np.arange(0, 256, 7), indexing="ij")
    px = np.stack([r, g, b], -1).astype(np.uint8).reshape(-1, 3)
    img = Image({"width": px.shape[0], "height": 1, "pixel": "rgb", "data": np.concatenate([px.reshape(-1), np.zeros(3, np.uint8)])})
    for to, wts in itertools.product(("grey", "r16"), (None, (0.2126, 0.7152, 0.0722))):
        opts = {"pixel": to}
        ow = None
        if wts:
            opts.update(redWeight=wts[0], greenWeight=wts[1], blueWeight=wts[2])
            ow = O.resolve_color_settings(*wts)
        want, ws = O.color_convert(np.ascontiguousarray(img.data), img.stride, img.width, 1, "rgb", to, ow)
        assert np.array_equal(P.colorConvertSync(img, opts).rows(), O.payload(want, ws, img.width, 1, to)), (to, wts)


def test_color_convert_1080p_cfg2(gpu):
    """BASELINE cfg2: 1080p rgba -> rgb / grey / greya, bit-exact on every payload byte."""
    P = gpu
    w, h = 1920, 1080
    img = Image({"width": w, "height": h, "pixel": "rgba", "data": fill_host(w, h, 4, w * 4, 1236, 0)})
    for to in ("rgb", "grey", "greya"):
        want, ws = O.color_convert(img.data, img.stride, w, h, "rgba", to)
        got = P.colorConvertSync(img, {"pixel": to})
        assert np.array_equal(got.rows(), O.payload(want, ws, w, h, to)), to


def test_cmyk_to_rgb_bit_exact(gpu):
    """The JPEG decoder's CMYK row loop (src/jpegcodec.cc:36-42): every (channel, K) pair, a wide image that
    takes the coalesced kernel, an odd width with row padding, and an unaligned subView."""
    P = gpu
    c, k = np.meshgrid(np.arange(256), np.arange(256))
    full = np.zeros((256, 256, 4), np.uint8)
    full[..., 0] = c; full[..., 1] = 255 - c; full[..., 2] = (c * 7 + 3) % 256; full[..., 3] = k
    img = Image({"width": 256, "height": 256, "pixel": "rgba", "data": full.reshape(-1).copy()})
    want, ws = O.cmyk_to_rgb(np.ascontiguousarray(img.data), img.stride, 256, 256)
    got = P.cmykToRgbSync(img)
    assert got.pixel == "rgb" and np.array_equal(got.rows(), O.payload(want, ws, 256, 256, "rgb"))
    rng = np.random.default_rng(77)
    for (w, h, pad, off) in [(1921, 37, 8, 0), (130, 9, 0, 3), (5, 4, 4, 1)]:
        img = rand_image(rng, w, h, "rgba", pad=pad, offset=off)
        want, ws = O.cmyk_to_rgb(np.ascontiguousarray(img.data), img.stride, w, h)
        got = P.cmykToRgbSync(img)
        assert np.array_equal(got.rows(), O.payload(want, ws, w, h, "rgb")), (w, h, pad, off)
    with pytest.raises(N.PichaError):
        P.cmykToRgbSync(rand_image(rng, 8, 8, "rgb"))


def test_convert_leaves_padding_untouched_and_handles_subviews(gpu):
    P = gpu
    rng = np.random.default_rng(3)
    parent = rand_image(rng, 200, 40, "rgb")
    view = parent.subView(7, 3, 150, 30)          # byte offset 3*600+21: unaligned base, parent stride
    want, ws = O.color_convert(np.ascontiguousarray(view.data), view.stride, 150, 30, "rgb", "rgba")
    got = P.colorConvertSync(view, {"pixel": "rgba"})
    assert np.array_equal(got.rows(), O.payload(want, ws, 150, 30, "rgba"))
    # C-ABI with a pre-filled dst whose stride has padding: padding bytes must survive
    src = rand_image(rng, 33, 6, "rgba")
    dst_buf = np.full(6 * 40, 0xAB, np.uint8)
    s = N.CImage(src.data.ctypes.data, src.stride, 33, 6, 1)
    d = N.CImage(dst_buf.ctypes.data, 40, 33, 6, 2)   # grey, 33 payload + 7 padding
    assert N.lib.picha_b200_color_convert(ctypes.byref(s), ctypes.byref(d), *O.resolve_color_settings()) == 0
    rows = dst_buf.reshape(6, 40)
    assert (rows[:, 33:] == 0xAB).all()
    want, ws = O.color_convert(src.data, src.stride, 33, 6, "rgba", "grey")
    assert np.array_equal(rows[:, :33], O.payload(want, ws, 33, 6, "grey"))


# ---- resize -----------------------------------------------------------------------------------

SHAPES = [
    (64, 48, 16, 12, 1.0),     # 4x down: lanczos has 17 taps against a 16-row ring
    (61, 47, 17, 13, 1.0),     # ragged down
    (16, 12, 48, 40, 0.7),     # up
    (40, 20, 20, 10, 1.0),     # 2x: box ring aliasing
    (30, 30, 10, 10, 1.0),     # 3x: box ring aliasing
    (33, 21, 47, 9, 1.5),      # up in x, down in y
    (23, 17, 23, 17, 1.0),     # same size
    (5, 300, 5, 4, 1.0),       # 75x vertical
    (300, 5, 4, 5, 1.0),       # 75x horizontal
    (1, 1, 7, 5, 1.0),         # single source pixel
    (9, 9, 1, 1, 1.0),         # single destination pixel
    (50, 50, 100, 100, 0.7),   # README example / cfg1
    (12, 12, 5, 7, 2.5),
]


@pytest.mark.parametrize("filt", N.FILTERS)
def test_resize_all_filters_formats_shapes(gpu, filt):
    P = gpu
    rng = np.random.default_rng(500 + N.FILTERS.index(filt))
    for pixel, (sw, sh, dw, dh, fw) in itertools.product(PIXEL_NAMES, SHAPES):
        img = rand_image(rng, sw, sh, pixel, pad=4 if (sw + dh) % 2 else 0)
        want = oracle_resize(img, dw, dh, filt, fw)
        for mode in MODES:
            got = P.resizeSync(img, dict({"width": dw, "height": dh, "filter": filt, "filterScale": fw}, **mode))
            assert_resize_close(got, want, mode.get("exact", False), (filt, pixel, sw, sh, dw, dh, fw, mode))


@pytest.mark.parametrize("pixel", PIXEL_NAMES)
def test_resize_fast_kernel_multi_tile(gpu, pixel):
    """Medium images: several column tiles and row bands per image, so tile origins, halos and the
    TMA box starts of every pixel size are exercised (the throughput kernel is the default here)."""
    P = gpu
    rng = np.random.default_rng(900 + PIXEL_NAMES.index(pixel))
    for (sw, sh, dw, dh, filt, fw) in [(1000, 400, 251, 97, "lanczos", 1.0), (777, 333, 200, 111, "cubic", 0.7),
                                       (301, 203, 640, 410, "mitchel", 1.0), (640, 360, 640, 90, "triangle", 1.0),
                                       (500, 500, 125, 250, "box", 1.0), (400, 300, 533, 100, "catmulrom", 1.3)]:
        img = rand_image(rng, sw, sh, pixel)
        want = oracle_resize(img, dw, dh, filt, fw)
        got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
        assert_resize_close(got, want, False, (pixel, sw, sh, dw, dh, filt, fw))


def test_resize_kernel_variants(gpu, monkeypatch):
    """Shapes picked for the code paths of the downscaling (csrc/resize_down.cuh) and upscaling
    (csrc/resize_up.cuh) kernels: pruned end taps and shared weight rows (integer ratios), 4- and
    8-row horizontal groups (forced both ways), the flat horizontal pass of 1- and 3-channel pixels,
    several row bands and launches per image, upscales close to 1:1 (widest per-thread window) and
    with a 6-row vertical window, and outputs narrower than one tile."""
    P = gpu
    rng = np.random.default_rng(4242)
    down = [("rgb", 3072, 200, 1024, 100, "lanczos", 1.0), ("grey", 4000, 160, 1333, 80, "lanczos", 1.0),   # tiles filled to the brim
            ("r16g16b16", 2400, 150, 800, 75, "catmulrom", 1.0),
            ("rgba", 1024, 600, 256, 150, "lanczos", 1.0), ("rgb", 1200, 640, 160, 160, "cubic", 0.7),
            ("grey", 999, 777, 333, 111, "mitchel", 1.0), ("greya", 1280, 720, 427, 241, "catmulrom", 1.0),
            ("r16g16b16", 900, 500, 300, 250, "triangle", 1.0), ("r16g16b16a16", 800, 1200, 237, 300, "lanczos", 1.0),
            ("rgba", 2000, 3000, 250, 1500, "box", 1.0), ("r16", 700, 900, 100, 450, "cubic", 1.0),
            # integer ratios of 4-channel pixels (sliding-window horizontal pass): 2, 3, 4; narrow and wide filters;
            # widths that are not multiples of a block or a tile; 16-bit
            ("rgba", 1000, 333, 500, 111, "cubic", 0.7), ("rgba", 1002, 400, 334, 190, "lanczos", 1.0),
            ("rgba", 1028, 300, 257, 77, "cubic", 0.7), ("r16g16b16a16", 768, 500, 256, 250, "mitchel", 1.0),
            ("rgba", 512, 256, 128, 64, "triangle", 1.0), ("r16g16b16a16", 1200, 300, 300, 100, "box", 1.0),
            ("rgba", 640, 480, 320, 240, "catmulrom", 1.4)]
    for group in ("4", "8", None):
        if group is None:
            monkeypatch.delenv("PICHA_B200_DOWN_G", raising=False)
        else:
            monkeypatch.setenv("PICHA_B200_DOWN_G", group)
        for (pixel, sw, sh, dw, dh, filt, fw) in down:
            img = rand_image(rng, sw, sh, pixel)
            want = oracle_resize(img, dw, dh, filt, fw)
            got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
            assert_resize_close(got, want, False, ("down", group, pixel, sw, sh, dw, dh, filt, fw))
            k = P.last_resize_kernel()    # 6: the integer-ratio horizontal pass (4-channel pixels), always 4-row groups
            assert k == 6 or k == {"4": 3, "8": 4}.get(group, k)
            assert k in (3, 4, 6)
            if k == 6:                      # and the general horizontal pass on the same shape
                monkeypatch.setenv("PICHA_B200_NO_P2INT", "1")
                got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
                monkeypatch.delenv("PICHA_B200_NO_P2INT")
                assert P.last_resize_kernel() in (3, 4)
                assert_resize_close(got, want, False, ("down-general", group, pixel, sw, sh, dw, dh, filt, fw))
    up = [("rgba", 300, 200, 1500, 1700, "mitchel", 1.0), ("r16g16b16a16", 257, 400, 771, 1601, "catmulrom", 1.0),
          ("rgba", 500, 300, 520, 310, "cubic", 1.0), ("r16g16b16a16", 200, 150, 333, 999, "lanczos", 1.5),
          ("rgba", 64, 300, 150, 1800, "triangle", 1.0), ("rgba", 640, 480, 1280, 960, "box", 1.0),
          ("rgba", 320, 200, 700, 900, "mitchel", 1.25), ("r16g16b16a16", 300, 200, 450, 333, "lanczos", 1.2),   # 5-row windows
          ("rgb", 333, 222, 1001, 667, "mitchel", 1.0), ("r16g16b16", 250, 180, 521, 377, "cubic", 1.0),
          ("grey", 400, 300, 1203, 450, "catmulrom", 1.0), ("r16", 300, 200, 450, 901, "lanczos", 1.0),
          ("greya", 320, 240, 642, 481, "triangle", 1.0), ("r16g16", 256, 256, 1023, 300, "mitchel", 1.0)]
    for (pixel, sw, sh, dw, dh, filt, fw) in up:
        img = rand_image(rng, sw, sh, pixel)
        want = oracle_resize(img, dw, dh, filt, fw)
        got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
        assert_resize_close(got, want, False, ("up", pixel, sw, sh, dw, dh, filt, fw))
        assert P.last_resize_kernel() == 5
        monkeypatch.setenv("PICHA_B200_OLD_UP", "1")       # the generic kernel on the same shape
        got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
        monkeypatch.delenv("PICHA_B200_OLD_UP")
        assert P.last_resize_kernel() == 2
        assert_resize_close(got, want, False, ("up-generic", pixel, sw, sh, dw, dh, filt, fw))


def test_resize_vertical_up_horizontal_down(gpu, monkeypatch):
    """A vertical upscale combined with a horizontal downscale (or any upscale whose 4-column groups touch more than 8
    source pixels): the upscaling kernel's wide-window variant (kernel 7, csrc/resize_up.cuh WPX = 0) -- every channel
    count and depth, widths that are not multiples of a tile or of 4, windows up to the 64-pixel cap; past the cap, and
    with the variant switched off, the bit-exact kernel serves the shape."""
    P = gpu
    rng = np.random.default_rng(909)
    shapes = [("rgb", 3000, 200, 800, 600, "lanczos", 1.0), ("rgba", 2000, 150, 700, 450, "cubic", 1.0),
              ("grey", 1777, 130, 431, 391, "mitchel", 1.0), ("greya", 1500, 140, 333, 500, "catmulrom", 1.0),
              ("r16g16b16a16", 1200, 140, 301, 421, "lanczos", 1.0), ("r16g16b16", 1600, 131, 500, 300, "triangle", 1.0),
              ("r16", 2048, 200, 256, 401, "cubic", 0.7), ("r16g16", 900, 130, 450, 390, "box", 1.0),
              ("rgb", 1000, 300, 980, 700, "lanczos", 1.5), ("rgba", 4000, 150, 500, 160, "triangle", 1.0)]
    for (pixel, sw, sh, dw, dh, filt, fw) in shapes:
        img = rand_image(rng, sw, sh, pixel)
        want = oracle_resize(img, dw, dh, filt, fw)
        got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
        assert P.last_resize_kernel() == 7, (pixel, sw, sh, dw, dh, filt, P.last_resize_kernel())
        assert_resize_close(got, want, False, ("wide-up", pixel, sw, sh, dw, dh, filt, fw))
    img = rand_image(rng, 3000, 200, "rgb")
    want = oracle_resize(img, 800, 600, "lanczos", 1.0)
    monkeypatch.setenv("PICHA_B200_NO_WIDE_UP", "1")
    got = P.resizeSync(img, {"width": 800, "height": 600, "filter": "lanczos"})
    monkeypatch.delenv("PICHA_B200_NO_WIDE_UP")
    assert P.last_resize_kernel() == 1 and got.equalPixels(want)
    img = rand_image(rng, 6000, 140, "rgb")            # 15:1 lanczos: 4 columns span more than 64 source pixels
    want = oracle_resize(img, 400, 300, "lanczos", 1.0)
    got = P.resizeSync(img, {"width": 400, "height": 300, "filter": "lanczos"})
    assert P.last_resize_kernel() == 1 and got.equalPixels(want)


def test_resize_column_pass_bank_groups(gpu):
    """The downscaling kernel's horizontal pass by columns (pass2_cols) deals a tile's columns to lanes by the
    16-byte bank group their tap window starts in.  Ratios that put EVERY window into the same group (windows 32
    floats apart: 8:1 rgba, 32:1 grey, 32:3 rgb) overflow the slots of that group and take the fallback that fills
    the holes; ratios one float off spread evenly.  Tiles wider and narrower than a CTA has threads."""
    P = gpu
    rng = np.random.default_rng(777)
    shapes = [("rgba", 2048, 300, 256, 150, "cubic", 1.0), ("rgba", 4096, 120, 512, 60, "triangle", 1.0),
              ("grey", 6400, 200, 200, 100, "cubic", 0.7), ("rgb", 3200, 260, 300, 130, "cubic", 0.7),
              ("r16g16b16a16", 2048, 200, 256, 50, "mitchel", 1.0), ("r16", 3200, 128, 100, 64, "triangle", 1.0),
              ("rgb", 3210, 260, 300, 130, "cubic", 0.7), ("grey", 1000, 400, 700, 100, "lanczos", 1.0),
              ("rgb", 1900, 300, 1300, 150, "mitchel", 1.0), ("rgba", 1500, 300, 1100, 140, "catmulrom", 1.0)]
    for (pixel, sw, sh, dw, dh, filt, fw) in shapes:
        img = rand_image(rng, sw, sh, pixel)
        want = oracle_resize(img, dw, dh, filt, fw)
        got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
        assert P.last_resize_kernel() == 3, (pixel, sw, sh, dw, dh, P.last_resize_kernel())
        assert_resize_close(got, want, False, ("cols", pixel, sw, sh, dw, dh, filt, fw))


def test_resize_random_shapes(gpu):
    """tests/fuzz_parity.py with a fixed seed: random formats, filters, filter scales, strides and independent
    x / y ratios between 1:5 up and 9:1 down, default path, against the oracle."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_parity.py"), "120", "11"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_resize_ill_conditioned_filters_take_the_exact_kernel(gpu):
    """lanczos / catmulrom narrowed to filterScale 0.5 have upscaling phases whose taps nearly cancel, and
    makeContribs (src/resize.cc:41-47) normalises them into weights of +-hundreds, +-thousands or non-finite
    values.  Nothing but the reference's own summation order reproduces those outputs, so such plans must be
    served by the bit-exact kernel even at sizes the throughput kernels would otherwise take (found by
    tests/fuzz_parity.py in wild mode)."""
    import picha_b200 as P
    rng = np.random.default_rng(77)
    for pixel, sw, sh, dw, dh, filt, fw in [("r16g16b16a16", 588, 1301, 187, 2849, "catmulrom", 0.5),
                                            ("greya", 804, 777, 46, 2919, "catmulrom", 0.5),
                                            ("r16g16b16", 1500, 1305, 6888, 177, "lanczos", 0.5)]:
        img = rand_image(rng, sw, sh, pixel)
        got = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt, "filterScale": fw})
        assert P.last_resize_kernel() == 1, (pixel, filt, fw)
        assert_resize_close(got, oracle_resize(img, dw, dh, filt, fw), True, (pixel, sw, sh, dw, dh, filt, fw))
    # the same filters at their usual widths are well conditioned and keep the throughput kernels
    img = rand_image(rng, 588, 1301, "r16g16b16a16")
    P.resizeSync(img, {"width": 187, "height": 420, "filter": "catmulrom"})
    assert P.last_resize_kernel() != 1
    # (a vertical upscale with a horizontal downscale: the upscaling kernel's wide-window variant)
    got = P.resizeSync(img, {"width": 187, "height": 2849, "filter": "catmulrom"})
    assert P.last_resize_kernel() == 7
    assert_resize_close(got, oracle_resize(img, 187, 2849, "catmulrom", 1.0), False, "wide-up, well conditioned")


def axis_matrix(filt, fw, src, dst, vertical):
    """The reference's filter along one axis as a dense float64 (dst x src) matrix, from the product's
    own contribution table; the vertical one uses the effective (ring-aliased) rows."""
    f = N.FILTERS.index(filt)
    ip, fp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)
    left = np.zeros(dst, np.int32); count = np.zeros(dst, np.int32); off = np.zeros(dst, np.int32)
    n = N.lib.picha_b200_contribs(f, fw, src, dst, left.ctypes.data_as(ip), count.ctypes.data_as(ip), off.ctypes.data_as(ip), None, None, 0)
    w = np.zeros(n, np.float32); eff = np.zeros(n, np.int32)
    N.lib.picha_b200_contribs(f, fw, src, dst, left.ctypes.data_as(ip), count.ctypes.data_as(ip), off.ctypes.data_as(ip),
                              w.ctypes.data_as(fp), eff.ctypes.data_as(ip), n)
    m = np.zeros((dst, src), np.float64)
    for i in range(dst):
        for k in range(count[i]):
            m[i, eff[off[i] + k] if vertical else left[i] + k] += float(w[off[i] + k])
    return m


def real_valued_resize(arr, dw, dh, filt, fw):
    """(dh, dw, C) float64 result of the reference's separable filter without any rounding."""
    h, w, _ = arr.shape
    u = arr.astype(np.float64) / 255.0
    mx, my = axis_matrix(filt, fw, w, dw, False), axis_matrix(filt, fw, h, dh, True)
    return np.einsum("yr,rxc->yxc", my, np.einsum("xs,rsc->rxc", mx, u))


def test_resize_extreme_ratios(gpu):
    """300:1 in either direction: ~1200-row (or -column) tap windows.  Only the bit-exact kernel takes these
    (narrow, one-row tiles), whatever mode is asked for."""
    P = gpu
    rng = np.random.default_rng(61)
    for pixel, (sw, sh, dw, dh) in itertools.product(("rgb", "r16g16b16a16"), [(64, 6000, 16, 20), (6000, 8, 20, 8)]):
        img = rand_image(rng, sw, sh, pixel)
        want = oracle_resize(img, dw, dh, "lanczos", 1.0)
        for mode in ({}, {"fast": True}):
            got = P.resizeSync(img, dict({"width": dw, "height": dh, "filter": "lanczos"}, **mode))
            assert got.equalPixels(want), (pixel, sw, sh, dw, dh, mode)


def test_resize_structured_inputs(gpu):
    """Constant rows, ramps, 0/max extremes, impulse (SURVEY 8d parity set)."""
    P = gpu
    w, h = 96, 64
    cases = {}
    cases["zeros"] = np.zeros((h, w, 4), np.uint8)
    cases["max"] = np.full((h, w, 4), 255, np.uint8)
    ramp = np.zeros((h, w, 4), np.uint8); ramp[:] = (np.arange(w) * 255 // (w - 1))[None, :, None]
    cases["ramp"] = ramp
    imp = np.zeros((h, w, 4), np.uint8); imp[31, 47] = 255
    cases["impulse"] = imp
    chk = np.zeros((h, w, 4), np.uint8); chk[::2, ::2] = 255; chk[1::2, 1::2] = 255
    cases["checker"] = chk
    for name, arr in cases.items():
        img = Image({"width": w, "height": h, "pixel": "rgba", "data": arr.reshape(-1).copy()})
        for filt, (dw, dh) in itertools.product(("lanczos", "cubic", "box", "mitchel"), [(24, 16), (33, 21), (192, 128)]):
            want = oracle_resize(img, dw, dh, filt, 1.0)
            truth = real_valued_resize(arr, dw, dh, filt, 1.0)
            for mode in MODES:
                got = P.resizeSync(img, dict({"width": dw, "height": dh, "filter": filt}, **mode))
                if mode.get("exact"):
                    assert_resize_close(got, want, True, (name, filt, dw, dh, mode))
                    continue
                # These images put many results exactly on a rounding tie (x.5), where the reference's own
                # choice is decided by its float rounding noise: any implementation that is not bit-identical
                # may land on the other side, so the mean bound does not apply.  What must hold: at most one
                # step from the reference, and a correct rounding of the real-valued result.
                g = got.rows().astype(np.float64).reshape(dh, dw, 4)
                assert np.abs(g - want.rows().astype(np.float64).reshape(dh, dw, 4)).max() <= 1, (name, filt, dw, dh, mode)
                assert np.abs(g - np.clip(truth * 255.0, 0, 255)).max() <= 0.5 + 1e-2, (name, filt, dw, dh, mode)


def test_resize_subview_and_padding(gpu):
    P = gpu
    rng = np.random.default_rng(11)
    parent = rand_image(rng, 120, 90, "rgb")
    view = parent.subView(5, 7, 100, 64)
    want = oracle_resize(view, 31, 19, "cubic", 0.7)
    assert_resize_close(P.resizeSync(view, {"width": 31, "height": 19, "exact": True}), want, True, "subview")
    assert_resize_close(P.resizeSync(view, {"width": 31, "height": 19}), want, False, "subview")
    assert_resize_close(P.resizeSync(view, {"width": 31, "height": 19, "fast": True}), want, False, "subview fast")
    dst_buf = np.full(19 * 100, 0xCD, np.uint8)
    s = N.CImage(view.data.ctypes.data, view.stride, 100, 64, 0)
    d = N.CImage(dst_buf.ctypes.data, 100, 31, 19, 0)
    assert N.lib.picha_b200_resize(ctypes.byref(s), ctypes.byref(d), 0, np.float32(0.7)) == 0
    rows = dst_buf.reshape(19, 100)
    assert (rows[:, 93:] == 0xCD).all()
    assert np.abs(rows[:, :93].astype(int) - want.rows().astype(int)).max() <= 1


@pytest.mark.parametrize("cfg", ["cfg3", "cfg4", "cfg5"])
def test_resize_benchmark_shapes(gpu, cfg):
    """One synthetic image of each BASELINE resize shape against the oracle (full size)."""
    P = gpu
    sw, sh, dw, dh, pixel, filt, opts, seed = {
        "cfg3": (3840, 2160, 960, 540, "rgba", "lanczos", {"filter": "lanczos"}, 1237),
        "cfg4": (2048, 2048, 4096, 4096, "r16g16b16a16", "mitchel", {"filter": "mitchel"}, 1238),
        "cfg5": (1920, 1080, 256, 256, "rgb", "cubic", {}, 1239),
    }[cfg]
    bpp = O.PIXEL_BYTES[O.PIXELS.index(pixel)]
    img = Image({"width": sw, "height": sh, "pixel": pixel, "data": fill_host(sw, sh, bpp, sw * bpp, seed, 0)})
    fw = 1.0 if "filter" in opts else 0.70
    want = oracle_resize(img, dw, dh, filt, fw)
    got = P.resizeSync(img, dict(opts, width=dw, height=dh))
    # the kernels the benchmark numbers are about: downscaling (4- / 8-row groups) and upscaling
    assert P.last_resize_kernel() == {"cfg3": 6, "cfg4": 5, "cfg5": 3}[cfg]
    assert_resize_close(got, want, False, cfg)
    if cfg != "cfg4":
        assert_resize_close(P.resizeSync(img, dict(opts, width=dw, height=dh, exact=True)), want, True, cfg)


def test_resize_properties_at_full_size(gpu):
    """Size-independent properties at BASELINE's full shapes: a constant image stays constant
    (weights are normalised), and resizing is deterministic call to call."""
    P = gpu
    for pixel, sw, sh, dw, dh, filt, val in [("rgba", 3840, 2160, 960, 540, "lanczos", 200),
                                             ("rgb", 1920, 1080, 256, 256, "cubic", 37)]:
        img = Image({"width": sw, "height": sh, "pixel": pixel})
        img.data[:] = val
        out = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt})
        assert np.abs(out.rows().astype(int) - val).max() <= 1
        again = P.resizeSync(img, {"width": dw, "height": dh, "filter": filt})
        assert out.equalPixels(again)


# ---- batches, device-resident entry points, threads -------------------------------------------

def test_batch_matches_single_calls(gpu):
    P = gpu
    rng = np.random.default_rng(21)
    imgs = [rand_image(rng, 160, 120, "rgba") for _ in range(9)]
    singles = [P.resizeSync(im, {"width": 40, "height": 30, "filter": "lanczos"}) for im in imgs]
    for device in (0, -1):
        batch = P.resizeBatchSync(imgs, {"width": 40, "height": 30, "filter": "lanczos"}, device=device)
        assert all(a.equalPixels(b) for a, b in zip(singles, batch))
    cs = [P.colorConvertSync(im, {"pixel": "grey"}) for im in imgs]
    cb = P.colorConvertBatchSync(imgs, {"pixel": "grey"}, device=-1)
    assert all(a.equalPixels(b) for a, b in zip(cs, cb))
    assert P.resizeBatchSync([], {"width": 4, "height": 4}) == []


def test_resize_convert_fused_equals_the_two_reference_calls(gpu):
    """picha_b200_resize_convert = doColorConvert(resizeImage(src)) (src/colorconvert.cc:171-188 on the output of
    src/resize.cc:270-280) in one kernel.  Exact mode: identical to the reference's two calls, every format pair.
    Throughput kernels: identical to this library's own resize followed by its (bit-exact) conversion, and within
    the resize tolerance of the reference pair where the conversion is monotone per channel."""
    P = gpu
    rng = np.random.default_rng(99)
    launches = []
    for sp in PIXEL_NAMES:
        if sp == "r16b16":
            continue
        for dp in PIXEL_NAMES:
            if dp == "r16b16":
                continue
            img = rand_image(rng, 61, 47, sp)
            ref = oracle_resize(img, 33, 29, "cubic", 0.7)
            want, ws = O.color_convert(np.ascontiguousarray(ref.data), ref.stride, 33, 29, sp, dp)
            got = P.resizeConvertSync(img, {"width": 33, "height": 29, "pixel": dp, "exact": True})
            assert np.array_equal(got.rows(), O.payload(want, ws, 33, 29, dp)), (sp, dp)
    # throughput kernels: downscale (general and integer-ratio passes), upscale, 8- and 16-bit, custom luma weights
    cases = [("rgba", "rgb", 1024, 600, 256, 150, "lanczos"), ("rgba", "grey", 1000, 333, 500, 111, "cubic"),
             ("rgb", "greya", 1200, 640, 160, 160, None), ("r16g16b16a16", "rgb", 800, 1200, 237, 300, "lanczos"),
             ("rgb", "r16g16b16a16", 900, 500, 300, 250, "triangle"), ("greya", "rgba", 1280, 720, 427, 241, "catmulrom"),
             ("rgba", "greya", 300, 200, 1500, 1700, "mitchel"), ("r16g16b16a16", "grey", 257, 400, 771, 1601, "catmulrom"),
             ("rgb", "rgba", 333, 222, 1001, 667, "mitchel"), ("grey", "rgb", 400, 300, 1203, 450, "catmulrom")]
    for sp, dp, sw, sh, dw, dh, filt in cases:
        img = rand_image(rng, sw, sh, sp)
        opts = {"width": dw, "height": dh, "pixel": dp, "redWeight": 0.3, "greenWeight": 0.5, "blueWeight": 0.2}
        if filt:
            opts["filter"] = filt
        before = P.launch_count()
        got = P.resizeConvertSync(img, opts)
        launches.append(P.launch_count() - before)
        k = P.last_resize_kernel()
        assert k in (3, 4, 5, 6), (sp, dp, k)
        two = P.colorConvertSync(P.resizeSync(img, opts), opts)
        assert got.equalPixels(two), (sp, dp, "fused differs from resize + convert", k)


def test_device_image_chain_never_leaves_the_device(gpu):
    """DeviceImage = lib/image.js's Image on HBM-resident pixels: resize -> subView -> colorConvert, copy
    and row on the device give what the reference chain gives on the host (README.md:33-37, test/copy.js:18-22)."""
    import torch
    from picha_b200.device import DeviceImage
    P = gpu
    rng = np.random.default_rng(123)
    host = rand_image(rng, 640, 480, "rgba", pad=12)
    dimg = DeviceImage.from_host(host)
    assert dimg.to_host().equalPixels(host)
    # the chain on the device, exact kernels, against the compiled reference on the host
    small = dimg.resize({"width": 200, "height": 150, "filter": "lanczos", "exact": True})
    view = small.subView(17, 9, 101, 77)                         # odd offsets: unaligned base, parent's stride
    grey = view.colorConvert({"pixel": "greya"})
    ref_small = oracle_resize(host, 200, 150, "lanczos", 1.0)
    ref_view = ref_small.subView(17, 9, 101, 77)
    want, ws = O.color_convert(np.ascontiguousarray(ref_view.data), ref_view.stride, 101, 77, "rgba", "greya")
    assert np.array_equal(grey.to_host().rows(), O.payload(want, ws, 101, 77, "greya"))
    assert view.to_host().equalPixels(ref_view)
    assert torch.equal(view.row(5).cpu(), torch.from_numpy(ref_view.row(5).copy()))
    # copy: test/copy.js:18-22 on the device
    target = DeviceImage({"width": 30, "height": 30, "pixel": "rgba"})
    small.copy(target)
    assert target.equalPixels(small.subView(0, 0, 30, 30))
    # throughput kernels through the handle: same tolerance as the host API; fused call = the two calls
    big = dimg.resize({"width": 160, "height": 120})
    assert_resize_close(big.to_host(), oracle_resize(host, 160, 120, "cubic", 0.7), False, "device image resize")
    fused = dimg.resizeConvert({"width": 160, "height": 120, "pixel": "grey"})
    assert fused.equalPixels(big.colorConvert({"pixel": "grey"}))


def test_batch_chunks_same_shape_runs_into_single_launches(gpu):
    """The host batch path cuts runs of same-shape images into chunks of one launch each; shapes, strides and pinned /
    pageable buffers may change from image to image.  Results equal the single calls; launches are far fewer than images."""
    P = gpu
    rng = np.random.default_rng(77)
    shapes = [(300, 200, "rgba", 0)] * 11 + [(301, 200, "rgba", 0)] + [(300, 200, "rgba", 8)] * 14 + [(300, 200, "rgb", 0)] * 9 + \
             [(640, 360, "rgb", 4)] * 5 + [(300, 200, "rgba", 0)] * 3
    imgs = [rand_image(rng, w, h, px, pad=pad) for (w, h, px, pad) in shapes]
    pins = []
    try:
        for i in (2, 3, 20, 30):          # a few sources in pinned memory, in the middle of runs
            im = imgs[i]
            ptr = N.lib.picha_b200_host_alloc(im.data.size)
            assert ptr
            pins.append(ptr)
            buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_ubyte)), shape=(im.data.size,))
            buf[:] = im.data
            imgs[i] = Image({"width": im.width, "height": im.height, "pixel": im.pixel, "stride": im.stride, "data": buf})
        opts = {"width": 75, "height": 50, "filter": "mitchel"}
        singles = [P.resizeSync(im, opts) for im in imgs]
        for device in (0, -1):
            before = P.launch_count()
            batch = P.resizeBatchSync(imgs, opts, device=device)
            launches = P.launch_count() - before
            assert all(a.equalPixels(b) for a, b in zip(singles, batch))
            if device == 0:
                assert launches <= 2 * 24, launches      # 43 images: at most 24 chunks (a descriptor write + a kernel each)
        grey = [P.colorConvertSync(im, {"pixel": "greya"}) for im in imgs]
        gb = P.colorConvertBatchSync(imgs, {"pixel": "greya"}, device=0)
        assert all(a.equalPixels(b) for a, b in zip(grey, gb))
    finally:
        for ptr in pins:
            N.lib.picha_b200_host_free(ptr)


def test_first_calls_of_many_shapes_from_many_threads(gpu):
    """Plans are built outside the device lock: 16 threads, 16 shapes nobody has asked for yet, all at once."""
    P = gpu
    rng = np.random.default_rng(78)
    imgs = [rand_image(rng, 211 + 7 * i, 157 + 5 * i, "rgba") for i in range(16)]
    got, errs = [None] * 16, []

    def work(i):
        try:
            got[i] = P.resizeSync(imgs[i], {"width": 53 + i, "height": 41 + i, "filter": "catmulrom"})
        except Exception as e:   # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    for i in range(16):
        want = oracle_resize(imgs[i], 53 + i, 41 + i, "catmulrom", 1.0)
        assert_resize_close(got[i], want, False, ("threads", i))


def test_device_resident_batch_and_synthetic_parity(gpu):
    import torch
    from picha_b200 import device as D
    n, sw, sh, dw, dh = 5, 512, 256, 128, 64
    src = D.DeviceBatch(n, sw, sh, "rgba")
    src.fill_synthetic(1237, first_image=3)
    torch.cuda.synchronize()
    for i in (0, 4):
        host = fill_host(sw, sh, 4, src.stride, 1237, 3 + i)
        assert np.array_equal(src.image(i).rows(), Image({"width": sw, "height": sh, "pixel": "rgba", "stride": src.stride, "data": host}).rows())
    dst = D.DeviceBatch(n, dw, dh, "rgba")
    D.resize(src, dst, "lanczos")
    grey = D.DeviceBatch(n, sw, sh, "greya")
    D.color_convert(src, grey)
    torch.cuda.synchronize()
    for i in range(n):
        im = src.image(i)
        assert_resize_close(dst.image(i), oracle_resize(im, dw, dh, "lanczos", 1.0), False, ("device", i))
        want, ws = O.color_convert(np.ascontiguousarray(im.data), im.stride, sw, sh, "rgba", "greya")
        assert np.array_equal(grey.image(i).rows(), O.payload(want, ws, sw, sh, "greya"))


def test_device_entry_with_unaligned_layout_uses_exact_kernel(gpu):
    """Caller-owned device images with odd strides cannot be TMA-staged: the call still succeeds (exact kernel)."""
    import torch
    from picha_b200 import device as D
    rng = np.random.default_rng(51)
    sw, sh, dw, dh = 300, 200, 75, 50
    img = rand_image(rng, sw, sh, "rgb")
    src = D.DeviceBatch(2, sw, sh, "rgb", stride=sw * 3 + 1)       # 901-byte rows
    dst = D.DeviceBatch(2, dw, dh, "rgb", stride=dw * 3 + 3)
    for i in range(2):
        src.upload(i, img)
    D.resize(src, dst, "lanczos")
    torch.cuda.synchronize()
    want = oracle_resize(img, dw, dh, "lanczos", 1.0)
    for i in range(2):
        assert np.array_equal(dst.image(i).rows(), want.rows())


def test_degenerate_and_error_cases_on_device(gpu):
    P = gpu
    empty = Image({"width": 0, "height": 3, "pixel": "rgb", "stride": 4, "data": np.zeros(12, np.uint8)})
    out = P.colorConvertSync(empty, {"pixel": "rgba"})
    assert out.width == 0 and out.height == 3
    one = Image({"width": 1, "height": 1, "pixel": "greya", "data": np.array([7, 200, 0, 0], np.uint8)})
    big = P.resizeSync(one, {"width": 3, "height": 2})
    assert (big.rows()[:, 0::2] == 7).all() and (big.rows()[:, 1::2] == 200).all()
    buf = np.zeros(64, np.uint8)
    a = N.CImage(buf.ctypes.data, 16, 4, 4, 1)
    b = N.CImage(buf.ctypes.data, 8, 2, 2, 0)
    assert N.lib.picha_b200_resize(ctypes.byref(a), ctypes.byref(b), 0, 1.0) == N.ERR_FORMAT_MISMATCH
    assert N.lib.picha_b200_color_convert(ctypes.byref(a), ctypes.byref(b), 0.3, 0.6, 0.1) == N.ERR_SIZE_MISMATCH
    assert N.lib.picha_b200_resize_batch(-1, None, None, 0, 1.0, 0, 0) == N.ERR_INVALID_ARGUMENT
    assert N.lib.picha_b200_init(99) == N.ERR_INVALID_ARGUMENT
    assert N.lib.picha_b200_init(0) == 0


def test_full_size_batch_properties(gpu):
    """cfg3 shape as a device-resident batch (size-independent properties): images with equal content give
    equal results wherever they sit in the batch, different content gives different results, a second
    run is bit-identical, and image 0 equals the single-image host call."""
    import torch
    from picha_b200 import device as D
    P = gpu
    n, sw, sh, dw, dh = 12, 3840, 2160, 960, 540
    src = D.DeviceBatch(n, sw, sh, "rgba")
    half = D.DeviceBatch(n // 2, sw, sh, "rgba")
    half.fill_synthetic(1237, first_image=40)
    src.buf[:half.buf.numel()] = half.buf                      # images 0..5
    src.buf[half.buf.numel():2 * half.buf.numel()] = half.buf  # images 6..11 = copies of 0..5
    dst = D.DeviceBatch(n, dw, dh, "rgba")
    D.resize(src, dst, "lanczos")
    torch.cuda.synchronize()
    first = dst.buf.clone()
    D.resize(src, dst, "lanczos")
    torch.cuda.synchronize()
    assert torch.equal(first, dst.buf)
    per = dst.buf[:n * dst.step].view(n, dst.step)
    for i in range(n // 2):
        assert torch.equal(per[i], per[i + n // 2]), i
    assert not torch.equal(per[0], per[1])
    host = P.resizeSync(src.image(0), {"width": dw, "height": dh, "filter": "lanczos"})
    assert host.equalPixels(dst.image(0))


def test_pinned_host_buffers(gpu):
    """Buffers from picha_b200_host_alloc take the no-staging path and give the same bytes."""
    P = gpu
    rng = np.random.default_rng(31)
    img = rand_image(rng, 200, 100, "rgba")
    n = img.data.size
    p = N.lib.picha_b200_host_alloc(n)
    q = N.lib.picha_b200_host_alloc(50 * 25 * 4)
    assert p and q
    try:
        pin_src = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_ubyte)), shape=(n,))
        pin_dst = np.ctypeslib.as_array(ctypes.cast(q, ctypes.POINTER(ctypes.c_ubyte)), shape=(50 * 25 * 4,))
        pin_src[:] = img.data
        s = N.CImage(p, img.stride, 200, 100, 1)
        d = N.CImage(q, 200, 50, 25, 1)
        assert N.lib.picha_b200_resize(ctypes.byref(s), ctypes.byref(d), 1, 1.0) == 0
        want = P.resizeSync(img, {"width": 50, "height": 25, "filter": "lanczos"})
        assert np.array_equal(pin_dst.reshape(25, 200), want.rows())
    finally:
        N.lib.picha_b200_host_free(p)
        N.lib.picha_b200_host_free(q)


def test_concurrent_calls_from_threads(gpu):
    """picha.resize runs on libuv pool threads, several at once (src/resize.cc:362-364)."""
    P = gpu
    rng = np.random.default_rng(41)
    imgs = [rand_image(rng, 97 + 3 * i, 61 + i, PIXEL_NAMES[i % 8]) for i in range(16)]
    want = [P.resizeSync(im, {"width": 31, "height": 17}) for im in imgs]
    got = [None] * len(imgs)
    errs = []

    def work(i):
        try:
            for _ in range(5):
                got[i] = P.resizeSync(imgs[i], {"width": 31, "height": 17})
        except Exception as e:   # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert all(a.equalPixels(b) for a, b in zip(want, got))
