"""Sharding of a batch of independent images across ranks / GPUs: contiguous blocks, no exchange.
Both functions ask the library (csrc/api.cu: shard_range, plan_chunks) -- the very code picha_b200_*_batch(...,
device=-1) runs -- so the multi-process tests on CPU exercise the product's own plumbing."""
import ctypes

from . import _native as N


def shard_range(n, rank, world):
    """Images [lo, hi) of shard `rank` of `world`."""
    lo, hi = ctypes.c_int(0), ctypes.c_int(0)
    if N.lib.picha_b200_shard_range(n, world, rank, ctypes.byref(lo), ctypes.byref(hi)) != 0:
        raise ValueError("bad rank/world")
    return lo.value, hi.value


def plan_batch(srcs, dsts, lanes=4):
    """[(first, count), ...]: the chunks of same-shape images (one kernel launch each) a batch on one device is cut
    into.  srcs / dsts: sequences of N.CImage (only width, height, pixel and stride are looked at)."""
    n = len(srcs)
    if n == 0:
        return []
    sa, da = (N.CImage * n)(*srcs), (N.CImage * n)(*dsts)
    first, count = (ctypes.c_int * n)(), (ctypes.c_int * n)()
    k = N.lib.picha_b200_plan_batch(n, sa, da, lanes, first, count, n)
    if k < 0:
        raise ValueError("bad batch")
    return [(first[i], count[i]) for i in range(k)]
