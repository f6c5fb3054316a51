"""Synthetic benchmark pixels: i.i.d. uniform bytes from a counter-based hash of
(seed, image, row, 32-bit word in the row payload).  ``fill_host`` reproduces on the CPU exactly
what picha_b200_synthetic_fill_device writes on the GPU (csrc/synthetic.cu), so the CPU baseline
and the parity checks can regenerate any image of a device-resident batch."""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def fill_host(width, height, bytes_per_pixel, stride, seed, image):
    """uint8 buffer of height*stride bytes; padding bytes are zero."""
    row_bytes = width * bytes_per_pixel
    words = (row_bytes + 3) // 4
    with np.errstate(over="ignore"):
        s = _mix64(np.array([seed], dtype=np.uint64))[0]
        y = np.arange(height, dtype=np.uint64)[:, None]
        k = np.arange(words, dtype=np.uint64)[None, :]
        ctr = (np.uint64(image) << np.uint64(40)) ^ (y << np.uint64(20)) ^ k
        v = (_mix64(s ^ ctr) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    payload = v.view(np.uint8).reshape(height, words * 4)[:, :row_bytes]
    out = np.zeros(height * stride, dtype=np.uint8)
    np.lib.stride_tricks.as_strided(out, shape=(height, row_bytes), strides=(stride, 1))[:] = payload
    return out
