import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import _native as N
torch.zeros(1, device="cuda")
buf = (C.c_longlong * 2)()
n = 64
for feat, name in ((0x00, "table only (per-iteration record)"), (0x10, "+ nk branches"), (0x20, "+ accumulate flag from record"),
                   (0x40, "+ LAST-flag exit"), (0x70, "all three")):
    iters = 200
    row0 = 0x80 | feat
    code = 1 | (1 << 1) | (80 << 8) | (row0 << 24)
    N.check_diag(N.diag_lib().drs_debug_mma_rate(n, iters, code, 1, buf))
    per = 36 if feat else 4
    print(f"{name:36s} N={n}: {buf[1] / (iters * per):7.1f} cyc/MMA")
