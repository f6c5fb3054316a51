"""Host-side logic and the C-ABI surface, CPU only (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O
import picha_b200 as P
from picha_b200 import _native as N
from picha_b200.image import Image
from picha_b200.shard import shard_range
from picha_b200.synthetic import fill_host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    """Every function include/picha_b200.h declares is exported by the library and bound in _native."""
    hdr = open(os.path.join(ROOT, "include", "picha_b200.h")).read()
    declared = set(re.findall(r"\b(picha_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared but not exported"
        assert name in N.SIGNATURES, f"{name} declared but not bound"
    assert set(N.SIGNATURES) == declared
    assert N.lib.picha_b200_version() == 100


def test_pixel_table_matches_reference_enum():
    # src/picha.h:79-92,118-172
    want = {"rgb": (0, 3, 3), "rgba": (1, 4, 4), "grey": (2, 1, 1), "greya": (3, 2, 2), "r16": (4, 2, 1),
            "r16g16": (5, 4, 2), "r16g16b16": (6, 6, 3), "r16g16b16a16": (7, 8, 4)}
    for name, (enum, nbytes, ch) in want.items():
        assert N.PIXELS[enum] == name
        assert N.lib.picha_b200_pixel_bytes(enum) == nbytes
        assert N.lib.picha_b200_pixel_channels(enum) == ch
        assert N.lib.picha_b200_row_stride(7, enum) == (nbytes * 7 + 3) & ~3
    assert N.lib.picha_b200_pixel_bytes(8) == 0 and N.lib.picha_b200_pixel_bytes(-1) == 0
    assert N.FILTERS == ["cubic", "lanczos", "catmulrom", "mitchel", "box", "triangle"]   # src/resize.cc:151-160


def test_resolve_resize_options():
    def res(has_f, tag, has_s, s):
        t, w = ctypes.c_int(-7), ctypes.c_float(-7)
        rc = N.lib.picha_b200_resolve_resize_options(has_f, tag, has_s, s, ctypes.byref(t), ctypes.byref(w))
        return rc, t.value, w.value
    assert res(0, 0, 0, 0.0) == (0, 0, np.float32(0.70))               # default cubic @ 0.70
    assert res(1, 1, 0, 0.0) == (0, 1, 1.0)                            # filter given -> width 1.0
    assert res(1, 5, 1, 1.5) == (0, 5, 1.5)
    assert res(0, 0, 1, 2.5) == (0, 0, 2.5)                            # filterScale alone keeps cubic
    assert res(1, 6, 0, 0.0)[0] == N.ERR_INVALID_FILTER
    assert res(1, -1, 0, 0.0)[0] == N.ERR_INVALID_FILTER
    assert res(0, 0, 1, float("nan"))[0] == N.ERR_INVALID_FILTER_WIDTH
    assert res(0, 0, 1, 0.0)[0] == N.ERR_INVALID_FILTER_WIDTH
    assert res(0, 0, 1, -1.0)[0] == N.ERR_INVALID_FILTER_WIDTH


def test_resolve_color_settings_matches_oracle():
    nan = float("nan")
    for args in [(nan, nan, nan), (0.2, 0.5, 0.3), (1, 1, 1), (nan, 2.0, nan), (3, nan, 1)]:
        out = (ctypes.c_float * 3)()
        N.lib.picha_b200_resolve_color_settings(*args, out)
        assert tuple(np.float32(v) for v in out) == tuple(np.float32(v) for v in O.resolve_color_settings(*args))
    out = (ctypes.c_float * 3)()
    N.lib.picha_b200_resolve_color_settings(nan, nan, nan, out)
    # the defaults sum to exactly 1.0f, so normalisation leaves them unchanged
    assert [float(v).hex() for v in out] == ["0x1.322d0e0000000p-2", "0x1.2c8b440000000p-1", "0x1.d2f1aa0000000p-4"]


@pytest.mark.parametrize("filt", range(6))
def test_contribs_equal_the_oracle_bit_for_bit(filt):
    shapes = [(3840, 960, 1.0), (2160, 540, 1.0), (1920, 256, 0.7), (1080, 256, 0.7), (2048, 4096, 1.0),
              (50, 100, 0.7), (76, 32, 0.7), (50, 24, 0.7), (100, 33, 1.3), (17, 17, 1.0), (9, 200, 0.5),
              (200, 9, 2.5), (31, 7, 1.0), (64, 16, 1.0), (30, 10, 1.0), (40, 20, 1.0)]
    for s, d, fw in shapes:
        l, r, o, w = O.contribs(filt, np.float32(fw), s, d)
        n = len(w)
        left = np.zeros(d, np.int32); count = np.zeros(d, np.int32); off = np.zeros(d, np.int32)
        wt = np.zeros(n, np.float32); eff = np.zeros(n, np.int32)
        ip = ctypes.POINTER(ctypes.c_int); fp = ctypes.POINTER(ctypes.c_float)
        got = N.lib.picha_b200_contribs(filt, fw, s, d, left.ctypes.data_as(ip), count.ctypes.data_as(ip),
                                        off.ctypes.data_as(ip), wt.ctypes.data_as(fp), eff.ctypes.data_as(ip), n)
        assert got == n, (filt, s, d, fw)
        assert np.array_equal(left, l) and np.array_equal(left + count - 1, r) and np.array_equal(off, o)
        assert np.array_equal(wt.view(np.uint32), w.view(np.uint32)), (filt, s, d, fw)
        # effective rows: simulate the reference's ring (src/resize.cc:83,99-108,126)
        scale = np.float32(s) / np.float32(d)
        ring_rows = {}
        centre = np.float32(0.5) * scale
        # fsupport as the table builder derives it; M from the widest possible window
        M = _ring_size(filt, fw, scale)
        srcrow = int(l.min()) if False else None
        filled = -1
        for y in range(d):
            need = _need_row(filt, fw, scale, centre, s)
            for row in range(filled + 1, need + 1):
                ring_rows[row % M] = row
            filled = max(filled, need)
            for q in range(count[y]):
                c = left[y] + q
                assert eff[off[y] + q] == ring_rows[c % M], (filt, s, d, fw, y, q)
            centre = np.float32(centre + scale)


def _support(filt, fw):
    base = {4: 0.5, 5: 1.0}.get(filt, 2.0)
    return np.float32(fw) * np.float32(base)


def _fsupport(filt, fw, scale):
    sup = _support(filt, fw)
    fscale = max(max(np.float32(scale), np.float32(1.0)), np.float32(1.0) / sup)
    return np.float32(sup * np.float32(fscale))


def _ring_size(filt, fw, scale):
    return int(np.ceil(np.float32(2) * _fsupport(filt, fw, scale)))


def _need_row(filt, fw, scale, centre, size):
    return min(size - 1, int(np.float32(centre + _fsupport(filt, fw, scale))))


def test_contribs_argument_errors():
    z = None
    assert N.lib.picha_b200_contribs(9, 1.0, 10, 10, z, z, z, z, z, 0) == N.ERR_INVALID_FILTER
    assert N.lib.picha_b200_contribs(0, 0.0, 10, 10, z, z, z, z, z, 0) == N.ERR_INVALID_FILTER_WIDTH
    assert N.lib.picha_b200_contribs(0, 1.0, 0, 10, z, z, z, z, z, 0) == N.ERR_INVALID_DIMENSIONS


# ---- Image semantics: lib/image.js -----------------------------------------------------------

def test_image_defaults_and_validation():
    im = Image({"width": 5, "height": 3, "pixel": "rgb"})
    assert im.stride == 16 and im.data.size == 48 and not im.data.any()
    assert Image().pixel == "rgba"
    with pytest.raises(ValueError, match="invalid pixel format"):
        Image({"width": 1, "height": 1, "pixel": "bgr"})
    with pytest.raises(ValueError, match="stride too short"):
        Image({"width": 4, "height": 1, "pixel": "rgba", "stride": 15})
    with pytest.raises(ValueError, match="image data too small"):
        Image({"width": 4, "height": 2, "pixel": "rgba", "data": np.zeros(31, np.uint8)})
    # lib/image.js:31 spells the 4-byte deep format 'r16b16' (SURVEY Q1); both names work here
    assert Image.pixelSize("r16b16") == 4 and Image.pixelSize("r16g16") == 4 and Image.pixelSize("nope") == 0


def test_copy_equals_subview(fixtures):
    """test/copy.js:18-22."""
    rows = fixtures["test2_jpg_rgb"]
    h, w, _ = rows.shape
    image = Image({"width": w, "height": h, "pixel": "rgb"})
    for y in range(h):
        image.row(y)[:] = rows[y].reshape(-1)
    c = Image({"width": 30, "height": 30, "pixel": "rgb"})
    image.copy(c)
    assert image.subView(0, 0, 30, 30).equalPixels(c)
    v = image.subView(3, 5, 20, 10)
    assert v.stride == image.stride and np.array_equal(v.row(2), image.row(7)[9:69])
    c.data[0] ^= 1
    assert not image.subView(0, 0, 30, 30).equalPixels(c)
    assert image.avgChannelDiff(Image({"width": w, "height": h, "pixel": "rgba"})) == 255


def test_metrics_ignore_padding():
    a = Image({"width": 3, "height": 2, "pixel": "rgb", "stride": 12})
    b = Image({"width": 3, "height": 2, "pixel": "rgb", "stride": 16})
    a.data[:] = 7
    b.data[:] = 9
    for y in range(2):
        b.row(y)[:] = 7
    assert a.equalPixels(b) and a.avgChannelDiff(b) == 0
    b.row(1)[4] = 10
    assert a.avgChannelDiff(b) == 3 / 18


# ---- argument / option errors of the public API (thrown before any device work) ----------------

def test_api_argument_errors_match_the_reference_messages():
    im = Image({"width": 4, "height": 4, "pixel": "rgb"})
    with pytest.raises(TypeError, match=r"expected: resizeSync\(image, opts\)"):
        P.resizeSync(im, None)
    with pytest.raises(TypeError, match=r"expected: resize\(image, opts, cb\)"):
        P.resize(im, {"width": 2, "height": 2}, None)
    with pytest.raises(TypeError, match=r"expected: colorConvertSync\(image, opts\)"):
        P.colorConvertSync(None, {})
    with pytest.raises(P.PichaError, match="invalid dimensions"):
        P.resizeSync(im, {"width": 2})
    with pytest.raises(P.PichaError, match="invalid dimensions"):
        P.resizeSync(im, {"width": -1, "height": 2})
    with pytest.raises(P.PichaError, match="invalid filter mode"):
        P.resizeSync(im, {"width": 2, "height": 2, "filter": "nearest"})
    with pytest.raises(P.PichaError, match="invalid filter width"):
        P.resizeSync(im, {"width": 2, "height": 2, "filterScale": 0})
    with pytest.raises(P.PichaError, match="invalid filter width"):
        P.resizeSync(im, {"width": 2, "height": 2, "filterScale": float("nan")})
    with pytest.raises(P.PichaError, match="expected pixel mode"):
        P.colorConvertSync(im, {})
    with pytest.raises(P.PichaError, match="invalid image"):
        P.resizeSync({"width": 4, "height": 4, "pixel": "rgb", "stride": 12, "data": np.zeros(10, np.uint8)},
                     {"width": 2, "height": 2})
    with pytest.raises(P.PichaError, match="invalid image"):
        P.colorConvertSync({"width": 4, "height": 0, "pixel": "rgb", "stride": 12, "data": np.zeros(48, np.uint8)},
                           {"pixel": "grey"})


def test_c_abi_validation_without_a_device():
    """Status codes come back before any CUDA call, so they are checkable on a CPU-only box."""
    buf = np.zeros(64, np.uint8)
    ok = N.CImage(buf.ctypes.data, 16, 4, 4, 1)
    small = N.CImage(buf.ctypes.data, 8, 2, 2, 1)
    other = N.CImage(buf.ctypes.data, 16, 4, 4, 0)
    null = N.CImage(None, 16, 4, 4, 1)
    badpix = N.CImage(buf.ctypes.data, 16, 4, 4, 9)
    short = N.CImage(buf.ctypes.data, 15, 4, 4, 1)
    R = N.lib.picha_b200_resize
    assert R(ctypes.byref(null), ctypes.byref(small), 0, 1.0) == N.ERR_INVALID_IMAGE
    assert R(ctypes.byref(badpix), ctypes.byref(small), 0, 1.0) == N.ERR_INVALID_IMAGE
    assert R(ctypes.byref(short), ctypes.byref(small), 0, 1.0) == N.ERR_INVALID_IMAGE
    assert R(ctypes.byref(ok), ctypes.byref(N.CImage(buf.ctypes.data, 8, 0, 2, 1)), 0, 1.0) == N.ERR_INVALID_DIMENSIONS
    assert R(ctypes.byref(ok), ctypes.byref(small), 6, 1.0) == N.ERR_INVALID_FILTER
    assert R(ctypes.byref(ok), ctypes.byref(small), 0, 0.0) == N.ERR_INVALID_FILTER_WIDTH
    assert R(ctypes.byref(ok), ctypes.byref(other), 0, 1.0) == N.ERR_FORMAT_MISMATCH
    C = N.lib.picha_b200_color_convert
    assert C(ctypes.byref(ok), ctypes.byref(small), 0.3, 0.6, 0.1) == N.ERR_SIZE_MISMATCH
    assert C(ctypes.byref(ok), ctypes.byref(badpix), 0.3, 0.6, 0.1) == N.ERR_INVALID_PIXEL
    assert N.lib.picha_b200_strerror(N.ERR_INVALID_IMAGE) == b"invalid image"
    if P.device_count() == 0:   # the product fails loudly instead of falling back to a CPU path
        assert R(ctypes.byref(ok), ctypes.byref(small), 0, 1.0) == N.ERR_NO_DEVICE
        assert C(ctypes.byref(ok), ctypes.byref(other), 0.3, 0.6, 0.1) == N.ERR_NO_DEVICE
        assert N.lib.picha_b200_host_alloc(16) is None


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under picha_b200/ may reference it."""
    pkg = os.path.join(ROOT, "picha_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "picha_oracle" not in text and "libpicha_ref" not in text, f


def test_shard_range_partitions_the_batch():
    for n in (0, 1, 7, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_host_generator_is_deterministic_and_uniform():
    a = fill_host(64, 32, 3, 196, 1237, 5)
    b = fill_host(64, 32, 3, 196, 1237, 5)
    c = fill_host(64, 32, 3, 196, 1237, 6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    rows = np.lib.stride_tricks.as_strided(a, (32, 192), (196, 1))
    assert abs(rows.mean() - 127.5) < 4 and rows.min() < 8 and rows.max() > 247
    pad = np.lib.stride_tricks.as_strided(a[192:], (31, 4), (196, 1))
    assert not pad.any()


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="no reference checkout here")
def test_addon_patch_applies_and_the_patched_sources_compile(tmp_path):
    """addon/picha_b200.patch turns the reference's four call sites (src/resize.cc:293,399,
    src/colorconvert.cc:201,288) and binding.gyp into calls of the C-ABI.  No Node toolchain exists here, so the
    check is: the patch applies cleanly to the reference's sources, and the WHOLE patched translation units compile
    (-fsyntax-only) against the inert v8/node/nan stand-ins the compiled oracle is built with."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    work = tmp_path / "picha"
    work.mkdir()
    shutil.copytree(os.path.join(REF, "src"), work / "src")
    shutil.copy(os.path.join(REF, "binding.gyp"), work / "binding.gyp")
    with open(os.path.join(root, "addon", "picha_b200.patch")) as f:
        r = subprocess.run(["patch", "-p1", "--batch"], cwd=work, stdin=f, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for unit in ("resize.cc", "colorconvert.cc"):
        text = (work / "src" / unit).read_text()
        assert "picha_b200::" in text and ("resizeImage(rsopts, src, dst);" not in text) and ("doColorConvert(cs, src, dst);" not in text)
        r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-I" + os.path.join(root, "oracle", "ref_shim"),
                            "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "addon"), "-I" + str(work / "src"),
                            str(work / "src" / unit)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    gyp = (work / "binding.gyp").read_text()
    assert "-lpicha_b200" in gyp and "picha_b200_root" in gyp


def test_batch_plan_chunks_same_shape_runs():
    """csrc/api.cu plan_chunks through the C-ABI: runs of same-shape images become chunks of one launch each, at
    least two chunks per lane when the batch allows, at most 32 images or ~128 MB of source per chunk."""
    from picha_b200.shard import plan_batch
    img = lambda w, h, p: N.CImage(1, w * N.lib.picha_b200_pixel_bytes(p), w, h, p)
    # 1024 thumbnails sources (1080p rgb, 6.2 MB each): 128 MB / 6.2 MB = 21 per chunk
    chunks = plan_batch([img(1920, 1080, 0)] * 1024, [img(256, 256, 0)] * 1024)
    assert sum(c for _, c in chunks) == 1024 and max(c for _, c in chunks) == 21 and len(chunks) == 49
    assert [f for f, _ in chunks] == sorted(f for f, _ in chunks) and chunks[0] == (0, 21)
    # 32 4K images (33 MB each) over 4 lanes: two chunks per lane = 4 each, which 128 MB just allows
    chunks = plan_batch([img(3840, 2160, 1)] * 32, [img(960, 540, 1)] * 32)
    assert chunks == [(4 * i, 4) for i in range(8)]
    # ... and 8K-wide images (133 MB each) go one by one
    assert max(c for _, c in plan_batch([img(7680, 4320, 1)] * 8, [img(960, 540, 1)] * 8, lanes=1)) == 1
    # small images: the count limit (at least 2 chunks per lane, at most 32 per chunk)
    assert max(c for _, c in plan_batch([img(64, 64, 1)] * 1000, [img(32, 32, 1)] * 1000)) == 32
    assert plan_batch([img(64, 64, 1)] * 16, [img(32, 32, 1)] * 16) == [(2 * i, 2) for i in range(8)]
    # a different shape (source or destination) ends a chunk
    srcs = [img(64, 64, 1)] * 5 + [img(65, 64, 1)] + [img(64, 64, 1)] * 4
    dsts = [img(32, 32, 1)] * 8 + [img(33, 32, 1)] * 2
    chunks = plan_batch(srcs, dsts, lanes=1)
    assert chunks == [(0, 5), (5, 1), (6, 2), (8, 2)]
    assert plan_batch([], []) == []


def test_wide_blocks_are_the_contribution_table_regrouped():
    """Host table of the upscaling kernel's wide-window variant (csrc/tables.cc: build_wide_blocks): block g lists, for
    the source pixels from column 4g's first tap on, the weight each carries into columns 4g .. 4g+3.  Regrouped by
    column it must be the reference's contribution table again (taps below 2^-30, which the fast-path tables prune,
    may be missing); axes whose blocks span at most 8 pixels, or more than 64, have no table."""
    import ctypes
    ip, fp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)
    for filt, fw, src, dst in [("lanczos", 1.0, 3000, 800), ("cubic", 0.7, 1777, 431), ("box", 1.3, 3745, 703),
                               ("triangle", 1.0, 999, 334), ("mitchel", 2.0, 640, 600)]:
        f = N.FILTERS.index(filt)
        window = N.lib.picha_b200_wide_blocks(f, fw, src, dst, None, 0)
        assert 8 < window <= 64, (filt, src, dst, window)
        groups = (dst + 3) // 4
        blocks = np.zeros(groups * window * 4, np.float32)
        assert N.lib.picha_b200_wide_blocks(f, fw, src, dst, blocks.ctypes.data_as(fp), blocks.size) == window
        blocks = blocks.reshape(groups, window, 4)
        left = np.zeros(dst, np.int32); count = np.zeros(dst, np.int32); off = np.zeros(dst, np.int32)
        n = N.lib.picha_b200_contribs(f, fw, src, dst, left.ctypes.data_as(ip), count.ctypes.data_as(ip), off.ctypes.data_as(ip), None, None, 0)
        w = np.zeros(n, np.float32)
        N.lib.picha_b200_contribs(f, fw, src, dst, left.ctypes.data_as(ip), count.ctypes.data_as(ip), off.ctypes.data_as(ip),
                                  w.ctypes.data_as(fp), None, n)
        dense = np.zeros((dst, src), np.float32)
        for x in range(dst):
            dense[x, left[x]:left[x] + count[x]] = w[off[x]:off[x] + count[x]]
        for g in range(groups):
            # the block starts at the first tap of column 4g that survives the pruning
            taps = np.nonzero(np.abs(dense[4 * g]) >= 2.0 ** -30)[0]
            lo = int(taps[0])
            for p in range(4):
                x = 4 * g + p
                got = np.zeros(src + window, np.float32)
                got[lo:lo + window] = blocks[g, :, p]
                if x >= dst:
                    assert not got.any()
                    continue
                want = np.where(np.abs(dense[x]) >= 2.0 ** -30, dense[x], 0)
                assert np.array_equal(got[:src], want), (filt, src, dst, x)
    # a plain 2x upscale keeps its blocks in registers (no table); a 20:1 lanczos downscale exceeds the window
    assert N.lib.picha_b200_wide_blocks(N.FILTERS.index("mitchel"), 1.0, 500, 1000, None, 0) == 0
    assert N.lib.picha_b200_wide_blocks(N.FILTERS.index("lanczos"), 1.0, 8000, 400, None, 0) == 0

