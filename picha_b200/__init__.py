"""picha_b200 -- B200-native implementation of jhs67/picha's pixel hot path.

The product is ``libpicha_b200.so`` (hand-written sm_100a kernels behind the C-ABI of
include/picha_b200.h).  This package is the host-side mirror of picha's JS surface for that
path -- ``Image`` (lib/image.js) and ``resize`` / ``resizeSync`` / ``colorConvert`` /
``colorConvertSync`` (index.js:13-33) -- plus the device-resident batch helpers the
benchmark uses.  Importing it requires the built library; there is no CPU fallback.
"""
from ._native import EXACT, FORCE_FAST, FILTERS, PIXELS, PichaError, lib   # noqa: F401
from .api import (cmykToRgbSync, colorConvert, colorConvertBatchSync, colorConvertSync, resize,   # noqa: F401
                  resizeBatchSync, resizeConvertSync, resizeSync)
from .image import Image   # noqa: F401


def device_count():
    return lib.picha_b200_device_count()


def launch_count():
    return lib.picha_b200_launch_count()


def last_resize_kernel():
    """1 bit-exact, 2 generic throughput, 3 / 4 downscaling (4- / 8-row groups), 5 upscaling kernel,
    6 downscaling with the integer-ratio horizontal pass (4-channel pixels at 2:1, 3:1, 4:1)."""
    return lib.picha_b200_last_resize_kernel()
