"""Diagnostic: one forward at a given (n, S) with per-layer errors against the oracle (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
from oracle import restatement as R
from test_gpu_unet import LAYERS

n, S = int(sys.argv[1]), int(sys.argv[2])
check = len(sys.argv) > 3
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres")
m.to(dev).eval()
x = T.np_randn(1, n, 3, S, S); lr = T.np_rand(2, 1, 3, S // 2, S // 2); t = torch.full((n,), 700)
with torch.no_grad():
    for it in range(3):
        got = m(x.to(dev), t.to(dev), lr.to(dev), 2)
        torch.cuda.synchronize()
        print("iter", it, "ok, finite:", bool(torch.isfinite(got).all()))
    if check:
        taps = {}
        ref = R.unet_forward(sd, "superres", x, t, lr, 2, None, taps)
        plan = m.native_plan(n, n, 1, S, 2)
        for name in LAYERS:
            have = m.debug_activation(plan, name, tuple(taps[name].shape))
            print(f"{name:8s} {T.max_rel_err(have, taps[name]):.3e}")
        print("eps", T.max_rel_err(got, ref))
