"""Diagnostic: tcgen05.mma rate by swizzle mode, K-slices per K-block, issuer count and descriptor source."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import _native as N
torch.zeros(1, device="cuda")
buf = (C.c_longlong * 2)()
iters = 1800
print("N nk layout issuers mode : issue cyc/MMA, complete cyc/MMA (per SM, all issuers)")
for n in (32, 64, 128, 256):
    for nk, layout in ((4, 2), (2, 4), (1, 6)):
        row16 = {2: 8, 4: 4, 6: 2}[layout]
        sbo16 = 18 * row16  # halo row pitch of an 8 x 16 tile with a 3 x 3 filter
        for issuers in ((1, 2) if n > 128 else (1, 2, 4)):
            for mode in (0, 1, 2):
                N.check_diag(N.diag_lib().drs_debug_mma_rate2(n, nk, layout, sbo16, issuers, iters, mode, buf))
                tot = iters * nk * issuers
                print(f"{n:4d} {nk} {layout} {issuers} {mode} : {buf[0] / (iters * nk):7.1f} {buf[1] / tot:7.1f}")
