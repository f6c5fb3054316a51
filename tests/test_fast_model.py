"""Host model of the throughput kernel (tools/fast_model.cc): the same vertical-axis tables
(accumulator ring / row window, csrc/tables.cc) and the same loop structure as csrc/resize_fast.cu,
run on the CPU and compared with the oracle.  It pins the table construction and every index the
kernel forms without needing a GPU; the GPU tests then only have to pin the CUDA translation."""
import ctypes
import itertools
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    so = os.path.join(ROOT, "tools", "libfast_model.so")
    srcs = [os.path.join(ROOT, "tools", "fast_model.cc"), os.path.join(ROOT, "picha_b200", "csrc", "tables.cc")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so] + srcs, check=True)
    lib = ctypes.CDLL(so)
    lib.fast_model.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    return lib


def run(model, rng, p, f, fw, sw, sh, dw, dh, band):
    ch, deep = O.PIXEL_CHANNELS[p], int(p >= 4)
    ss = O.row_stride(sw, p)
    src = rng.integers(0, 256, ss * sh, dtype=np.uint8)
    ds = O.row_stride(dw, p)
    dst = np.zeros(ds * dh, np.uint8)
    info = (ctypes.c_int * 6)()
    rc = model.fast_model(f, fw, src.ctypes.data, ss, sw, sh, dst.ctypes.data, ds, dw, dh, ch, deep, band, 0, info)
    if rc == 1:
        return None, list(info)        # shape outside the fast kernel's reach: the product uses the exact kernel
    want, _ = O.resize(src, ss, sw, sh, p, dw, dh, f, fw)
    d = np.abs(O.channels(dst, ds, dw, dh, p).astype(int) - O.channels(want, ds, dw, dh, p).astype(int))
    return d, list(info)


SHAPES = [(64, 48, 16, 12, 1.0), (61, 47, 17, 13, 1.0), (16, 12, 48, 40, 0.7), (40, 20, 20, 10, 1.0),
          (30, 30, 10, 10, 1.0), (33, 21, 47, 9, 1.5), (23, 17, 23, 17, 1.0), (5, 300, 5, 4, 1.0), (300, 5, 4, 5, 1.0),
          (1, 1, 7, 5, 1.0), (9, 9, 1, 1, 1.0), (50, 50, 100, 100, 0.7), (12, 12, 5, 7, 2.5), (200, 120, 50, 30, 1.0),
          (120, 90, 333, 200, 1.0)]


@pytest.mark.parametrize("filt", range(6))
def test_model_matches_oracle_within_tolerance(model, filt):
    rng = np.random.default_rng(40 + filt)
    covered = {0: 0, 1: 0}
    for p, (sw, sh, dw, dh, fw), band in itertools.product((0, 1, 2, 7), SHAPES, (8, 24)):
        d, info = run(model, rng, p, filt, fw, sw, sh, dw, dh, band)
        if d is None:
            continue
        assert info[3] == 0, ("index check failed", p, filt, sw, sh, dw, dh, fw, band, info)
        covered[info[0]] += 1
        assert d.max() <= 1, (p, filt, sw, sh, dw, dh, fw, band, int(d.max()))
        if d.size >= 4096:
            assert d.mean() <= 0.05
        else:
            assert int((d > 0).sum()) <= max(1, int(0.05 * d.size)), (p, filt, sw, sh, dw, dh, fw, band)
    assert covered[0] > 10 and covered[1] > 10     # both vertical forms were exercised


def test_model_benchmark_shapes_quarter_size(model):
    rng = np.random.default_rng(77)
    for p, f, fw, sw, sh, dw, dh, band, variant, depth in [(1, 1, 1.0, 960, 540, 240, 135, 24, 0, 4),     # cfg3 / 4 (end taps of ~1e-16 pruned: 16 taps, 4 open rows)
                                                             (0, 0, 0.7, 480, 270, 64, 64, 16, 0, 3),         # cfg5 / 4
                                                             (7, 3, 1.0, 256, 256, 512, 512, 112, 1, 4)]:    # cfg4 / 8
        d, info = run(model, rng, p, f, np.float32(fw), sw, sh, dw, dh, band)
        assert info[0] == variant and info[1] == depth and info[3] == 0, info
        assert d.max() <= 1 and d.mean() <= 0.05
