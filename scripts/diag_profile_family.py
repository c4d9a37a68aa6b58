"""Diagnostic: per-launch profile (one CUDA event per launch) of any family / batch / size.
usage (GPU box): python scripts/diag_profile_family.py <superres|sar|generation> <nb> <nx> <S> [mag]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
from diffusionremotesensing_b200 import _native as N
fam, nb, nx, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mag = int(sys.argv[5]) if len(sys.argv) > 5 else (2 if fam == "superres" else 1)
dev = torch.device("cuda:0")
m, sd = T.default_init_model(fam); m.to(dev).eval()
xc = 1 if fam == "sar" else 3
plan = m.native_plan(nb, nx, 1, S, mag)
x = T.np_randn(1, nx, xc, S, S).to(dev)
eps = torch.empty((nb, xc, S, S), device=dev)
st = N.stream_ptr(dev); lib = N.lib()
nl = lib.drs_plan_launch_count(plan)
ms = torch.zeros(nl)
for _ in range(2):
    N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 5, N.ptr(ms), st))
N.check(lib.drs_plan_check(plan, st))
nm = ctypes.create_string_buffer(64)
fl, by, ct, sm = ctypes.c_double(), ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
tot = 0.0
for i in range(nl):
    lib.drs_plan_launch_info(plan, i, nm, 64, ctypes.byref(fl), ctypes.byref(by), ctypes.byref(ct), ctypes.byref(sm))
    t = float(ms[i]) * 1e3
    tot += t
    print(f"{nm.value.decode():28s} {t:8.1f} us {fl.value / 1e9 / max(t, 1e-9) * 1e3:8.1f} TF/s  ctas {ct.value:6d} smem {sm.value // 1024:4d} KiB")
print(f"total {tot:.1f} us")
