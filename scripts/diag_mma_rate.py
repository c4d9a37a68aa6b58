import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import _native as N
torch.zeros(1, device="cuda")
buf = (C.c_longlong * 2)()
for hammer, name in ((0, "idle"), (1, "tcgen05.ld loop"), (2, "LDS loop"), (3, "STG loop")):
    for n in (32, 64, 128):
        iters = 2000
        code = 1 | (hammer << 1) | (80 << 8) | (11 << 24)
        N.check(N.lib().drs_debug_mma_rate(n, iters, code, 1, buf))
        print(f"other warps: {name:16s} N={n:3d}: {buf[1] / (iters * 4):7.1f} cyc/MMA")
