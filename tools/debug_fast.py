"""Debug helper: device-resident resize of one shape; prints the CUDA status.
usage: debug_fast.py pixel sw sh dw dh filter|none width n"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from picha_b200 import device as D

pixel, sw, sh, dw, dh, filt, width, n = sys.argv[1:9]
sw, sh, dw, dh, n, width = int(sw), int(sh), int(dw), int(dh), int(n), float(width)
src = D.DeviceBatch(n, sw, sh, pixel)
dst = D.DeviceBatch(n, dw, dh, pixel)
src.fill_synthetic(7)
torch.cuda.synchronize()
try:
    D.resize(src, dst, None if filt == "none" else filt, width)
    torch.cuda.synchronize()
    print("OK ", *sys.argv[1:9], int(dst.buf[:64].sum()))
except Exception as e:
    print("ERR", *sys.argv[1:9], str(e).splitlines()[0][:100])
