"""Stress single convolutions through drs_debug_conv2d (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from diffusionremotesensing_b200 import _native as N
from test_gpu_conv_layers import run_native, bf16r, ref_conv
dev = torch.device("cuda:0")
cases = [("3x3s2", 16, 64, 64, 128, 128), ("3x3", 16, 64, 64, 128, 128)] if len(sys.argv) > 2 else [("3x3", 16, 16, 32, 256, 256), ("3x3", 16, 32, 32, 256, 256), ("3x3s2", 16, 32, 32, 256, 256), ("3x3", 16, 32, 64, 128, 128),
         ("3x3", 16, 64, 64, 128, 128), ("3x3s2", 16, 64, 64, 128, 128), ("3x3", 16, 64, 128, 64, 64), ("3x3", 16, 128, 128, 64, 64),
         ("3x3", 16, 128, 256, 32, 32), ("3x3", 16, 256, 256, 32, 32), ("T3x3s2", 16, 256, 256, 32, 32), ("T3x3s2", 16, 128, 128, 64, 64),
         ("T3x3s2", 16, 64, 64, 128, 128), ("1x1", 16, 256, 128, 32, 32), ("1x1", 16, 32, 32, 256, 256), ("2x2s2", 16, 32, 32, 256, 256)]
reps = int(sys.argv[1])
for kind, B, Cin, Cout, H, W in cases:
    x = bf16r(torch.randn(B, Cin, H, W)).to(dev)
    wshape = (Cin, Cout, 3, 3) if kind == "T3x3s2" else (Cout, Cin) + {"3x3": (3, 3), "3x3s2": (3, 3), "1x1": (1, 1), "2x2s2": (2, 2)}[kind]
    w = bf16r(torch.randn(wshape) / (Cin * 3) ** 0.5).to(dev)
    b = torch.randn(Cout).to(dev)
    ok = True
    y0 = None
    try:
        for r in range(reps):
            y = run_native(x, w, b, None, None, kind, False)
            if y0 is None:
                y0 = y.clone()
            elif not torch.equal(y, y0):
                print(kind, Cin, Cout, H, "rep", r, "RESULT CHANGED max diff", float((y - y0).abs().max()))
                ok = False
                break
    except Exception as e:
        print(kind, Cin, Cout, H, "FAILED:", str(e)[:150])
        ok = False
        break
    print(kind, Cin, Cout, H, "ok" if ok else "BAD")
