"""Randomised check that the device resize writes nothing but the destination's payload bytes (run on a GPU
box): python tests/fuzz_canary.py [cases] [seed] [wild].  The destination sits in the middle of a canary-filled
buffer; every byte outside width*bytes of each row -- row padding, and 4 MiB either side -- must survive."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import picha_b200 as P
from picha_b200 import _native as N, device as D
from picha_b200.image import PIXEL_ENUM, PIXEL_NAMES, PIXEL_SIZES

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
wild = len(sys.argv) > 3 and sys.argv[3] == "wild"
MARGIN = 4 << 20
bad, served, skipped = 0, {}, 0
rng2 = np.random.default_rng(5)
for i in range(cases):
    pixel = PIXEL_NAMES[rng.integers(0, 8)]
    filt = N.FILTERS[rng.integers(0, 6)]
    fw = float(rng.choice([0.7, 1.0, 1.0, 1.3, 2.0] + ([0.5, 3.0] if wild else [])))
    sw, sh = int(rng.integers(130, 9000 if wild else 2600)), int(rng.integers(130, 1500 if wild else 700))
    lo, hi = (1 / 12.0, 40.0) if wild else (0.2, 9.0)
    rx, ry = float(np.exp(rng.uniform(np.log(lo), np.log(hi)))), float(np.exp(rng.uniform(np.log(lo), np.log(hi))))
    dw, dh = max(1, int(sw / rx)), max(1, int(sh / ry))
    if dw * dh > 6_000_000 or dw > 8000 or dh > 4000:
        continue
    bpp = PIXEL_SIZES[pixel]
    hstride = ((sw * bpp + 3) & ~3) + int(rng.choice([0, 0, 4, 12]))
    rng.integers(0, 256, hstride * sh, dtype=np.uint8)            # (keeps the case sequence of fuzz_parity.py)
    src = D.DeviceBatch(1, sw, sh, pixel)
    src.buf.random_(0, 256)
    dst = D.DeviceBatch(1, dw, dh, pixel, stride=(dw * bpp + 127) // 128 * 128 + int(rng2.choice([0, 0, 16, 48, 4])))
    span = dst.stride * dh
    dst.buf = torch.full((span + 2 * MARGIN,), 0xA5, dtype=torch.uint8, device=src.device)
    inner = dst.buf[MARGIN:MARGIN + span]
    dst.cimage = lambda index=0, d=dst, p=inner.data_ptr(): N.CImage(p, d.stride, d.width, d.height, PIXEL_ENUM[d.pixel])
    try:
        D.resize(src, dst, filt, fw)
        torch.cuda.synchronize()
    except N.PichaError as e:
        if "invalid" in str(e) or "unsupported" in str(e).lower():
            skipped += 1
            continue
        raise
    k = P.last_resize_kernel()
    served[k] = served.get(k, 0) + 1
    outside = int((dst.buf[:MARGIN] != 0xA5).sum()) + int((dst.buf[MARGIN + span:] != 0xA5).sum())
    pad = int((inner.view(dh, dst.stride)[:, dw * bpp:] != 0xA5).sum())
    if outside or pad:
        bad += 1
        print("STRAY WRITES", i, pixel, sw, sh, "->", dw, dh, filt, fw, "dst stride", dst.stride, "kernel", k,
              "outside", outside, "row padding", pad, flush=True)
print("canary cases by kernel:", dict(sorted(served.items())), "rejected:", skipped, "bad:", bad)
sys.exit(1 if bad else 0)
