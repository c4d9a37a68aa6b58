import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import _native as N
torch.zeros(1, device="cuda")
buf = (C.c_longlong * 2)()
for spin, name in ((0, "other warps parked at bar.sync"), (1, "other warps spin on the mbarrier")):
    for n in (32, 64, 128):
        iters = 1800
        code = 1 | (1 << 1) | (spin << 2) | (80 << 8) | (0x80 << 24)
        N.check(N.lib().drs_debug_mma_rate(n, iters, code, 1, buf))
        print(f"{name:34s} N={n:3d}: {buf[1] / (iters * 4):7.1f} cyc/MMA")
