import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import _native as N
torch.zeros(1, device="cuda")
buf = (C.c_longlong * 2)()
for ctas in (1, 2, 3):
    for unroll in (0, 1):
        for n in (16, 32, 64, 128, 256):
            iters = 2000
            N.check(N.lib().drs_debug_mma_rate(n, iters, unroll, ctas, buf))
            cnt = iters * (4 if unroll else 1)
            print(f"ctas/SM={ctas} unroll4={unroll} N={n:3d}: issue {buf[0] / cnt:7.1f} cyc/MMA, complete {buf[1] / cnt:7.1f} cyc/MMA (tensor floor {n / 2:.0f})")
