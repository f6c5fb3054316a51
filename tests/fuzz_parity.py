"""Randomised parity of the default resize path against the oracle (run on a GPU box):
python tests/fuzz_parity.py [cases] [seed].  Prints every case out of tolerance with the kernel that
served it and exits non-zero if there was one."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import picha_b200 as P
from picha_b200 import _native as N
from picha_b200.image import Image, PIXEL_NAMES

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
wild = len(sys.argv) > 3 and sys.argv[3] == "wild"      # wider ranges: sizes to 9000, ratios 1:12 .. 40:1, filter scales to 3
bad, unsupported, served = 0, 0, {}
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
    bpp = O.PIXEL_BYTES[O.PIXELS.index(pixel)]
    stride = ((sw * bpp + 3) & ~3) + int(rng.choice([0, 0, 4, 12]))
    img = Image({"width": sw, "height": sh, "pixel": pixel, "stride": stride,
                 "data": rng.integers(0, 256, stride * sh, dtype=np.uint8)})
    if i < int(os.environ.get("FUZZ_SKIP_BELOW", "0")) or str(i) in os.environ.get("FUZZ_SKIP_CASES", "").split(","):
        continue                      # (debugging: replay the random sequence without running the early cases)
    try:
        want, ws = O.resize(np.ascontiguousarray(img.data), stride, sw, sh, pixel, dw, dh, filt, fw)
    except Exception:
        continue                      # shapes the reference itself rejects (total weight 0)
    if os.environ.get("FUZZ_VERBOSE"):
        print("case", i, pixel, sw, sh, "->", dw, dh, filt, fw, "stride", stride, flush=True)
    opts = {"width": dw, "height": dh, "filter": filt, "filterScale": fw}
    if os.environ.get("FUZZ_EXACT_CASE") == str(i):
        opts["exact"] = True          # (debugging: take one case through the bit-exact kernel instead)
    try:
        got = P.resizeSync(img, opts)
    except P.PichaError as e:
        if "unsupported" not in str(e):
            raise                     # a CUDA fault ends the run (and the context)
        unsupported += 1
        print("UNSUPPORTED", pixel, sw, sh, "->", dw, dh, filt, fw, flush=True)
        continue
    k = P.last_resize_kernel()
    served[k] = served.get(k, 0) + 1
    a = np.ascontiguousarray(got.rows())
    b = np.ascontiguousarray(O.payload(want, ws, dw, dh, pixel))
    if bpp // O.PIXEL_CHANNELS[O.PIXELS.index(pixel)] == 2:
        a, b = a.view(np.uint16), b.view(np.uint16)
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    ok = d.max() <= 1 and (d.mean() <= 0.05 if d.size >= 4096 else (d > 0).sum() <= max(1, int(0.05 * d.size)))
    if not ok:
        bad += 1
        print("OUT OF TOLERANCE", pixel, sw, sh, "->", dw, dh, filt, fw, "stride", stride, "kernel", k, "max", int(d.max()), "mean", float(d.mean()))
print("cases by kernel (1 exact, 2 generic, 3/4 down, 5 up, 6 down with the integer-ratio pass, 7 up with the wide window):", dict(sorted(served.items())),
      "unsupported:", unsupported, "bad:", bad)
sys.exit(1 if bad else 0)
