"""Time one resize shape on device-resident synthetic images (not part of bench.py's contract):
python tools/bench_shape.py PIXEL SW SH DW DH FILTER [FILTER_SCALE] [N] -> ms per batch, GB/s of
algorithmic bytes, fraction of the measured HBM peak, and which kernel ran."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import picha_b200 as P
from picha_b200 import device as D
from picha_b200.image import PIXEL_SIZES

pixel, sw, sh, dw, dh, filt = sys.argv[1], *map(int, sys.argv[2:6]), sys.argv[6]
fs = float(sys.argv[7]) if len(sys.argv) > 7 and sys.argv[7] != "-" else None
n = int(sys.argv[8]) if len(sys.argv) > 8 else 16
src, dst = D.DeviceBatch(n, sw, sh, pixel), D.DeviceBatch(n, dw, dh, pixel)
src.fill_synthetic(99)
for _ in range(3):
    D.resize(src, dst, filt, fs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    D.resize(src, dst, filt, fs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_ = src.payload_bytes + dst.payload_bytes
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
print(f"{pixel} {sw}x{sh}->{dw}x{dh} {filt} x{n}: {ms:.3f} ms  {bytes_ / ms / 1e6:.0f} GB/s  frac {bytes_ / ms / 1e6 / peak:.3f}  kernel {P.last_resize_kernel()}")
