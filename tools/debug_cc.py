"""Debug helper: device-resident colour conversion of a 1080p batch. usage: debug_cc.py src dst n"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from picha_b200 import device as D
sp, dp, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
src = D.DeviceBatch(n, 1920, 1080, sp)
dst = D.DeviceBatch(n, 1920, 1080, dp)
src.fill_synthetic(7)
torch.cuda.synchronize()
for _ in range(2):
    D.color_convert(src, dst)
torch.cuda.synchronize()
print("OK", sp, dp, n, int(dst.buf[:64].sum()))
