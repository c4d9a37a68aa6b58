import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import _native as N
torch.zeros(1, device="cuda")
buf = (C.c_longlong * 2)()
for sbo, row0 in ((64, 0), (80, 0), (80, 1), (80, 11), (72, 3), (64, 3)):
    for n in (32, 64, 128):
        iters = 2000
        code = 1 | (sbo << 8) | (row0 << 24)
        N.check(N.lib().drs_debug_mma_rate(n, iters, code, 1, buf))
        cnt = iters * 4
        print(f"SBO={sbo * 16:5d} B start row {row0:2d} N={n:3d}: {buf[1] / cnt:7.1f} cyc/MMA")
