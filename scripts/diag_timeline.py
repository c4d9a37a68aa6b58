"""Diagnostic: timeline of CTA 0 of one debug conv launch (run with DRS_V2_TIMELINE=1 on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import torch
from diffusionremotesensing_b200 import _native as N
from test_gpu_conv_layers import run_native, bf16r

kind, B, Cin, Cout, H, W = sys.argv[1], *[int(v) for v in sys.argv[2:7]]
dev = torch.device("cuda:0")
x = bf16r(torch.randn(B, Cin, H, W)).to(dev)
wshape = (Cin, Cout, 3, 3) if kind == "T3x3s2" else (Cout, Cin) + {"3x3": (3, 3), "3x3s2": (3, 3), "1x1": (1, 1), "2x2s2": (2, 2)}[kind]
w = bf16r(torch.randn(wshape) / (Cin * 3) ** 0.5).to(dev)
b = torch.randn(Cout).to(dev)
for _ in range(2):
    y = run_native(x, w, b, None, None, kind, False)
buf = (C.c_longlong * 512)()
N.check(N.lib().drs_debug_timeline(buf, 512))
t0 = buf[0]
names = ["p_start", "p_issued", "m_tmemfree", "m_afull", "m_issued", "e_tfull", "e_done"]
print("tile " + " ".join(f"{n:>11s}" for n in names))
for t in range(12):
    print(f"{t:4d} " + " ".join(f"{buf[t * 8 + s] - t0:11d}" for s in range(7)))
print("K-block start stamps of tile 3 (issuing lane):", [buf[256 + kb] - buf[256] for kb in range(20) if buf[256 + kb] > 0])
