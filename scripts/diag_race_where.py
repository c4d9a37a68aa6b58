import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_conv_layers import run_native, bf16r, ref_conv
dev = torch.device("cuda:0")
kind, B, Cin, Cout, H, W = "3x3s2", 16, 64, 64, 128, 128
x = bf16r(torch.randn(B, Cin, H, W)).to(dev)
w = bf16r(torch.randn(Cout, Cin, 3, 3) / (Cin * 3) ** 0.5).to(dev)
b = torch.randn(Cout).to(dev)
ref = ref_conv(x.double(), w.double(), b.double(), kind)
for r in range(12):
    y = run_native(x, w, b, None, None, kind, False)
    bad = ((y.double() - ref).abs() > 0.05 * ref.abs().max()).any(dim=1)   # [B, OH, OW]
    if bad.any():
        idx = bad.nonzero()
        tiles = sorted({(int(i[0]), int(i[1]) // 16, int(i[2]) // 8) for i in idx})
        lin = sorted({t[0] * 32 + t[1] * 8 + t[2] for t in tiles})
        print("rep", r, "bad pixels", int(bad.sum()), "tiles(b,ty,tx)", tiles[:8], "linear tile ids", lin[:12], "ids % 148:", sorted({l % 148 for l in lin})[:12])
    else:
        print("rep", r, "all good")
