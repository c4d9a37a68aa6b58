"""Diagnostic: forwards back to back without host syncs (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
from diffusionremotesensing_b200 import _native as N
n, S, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
x = T.np_randn(1, n, 3, S, S).to(dev); lr = T.np_rand(2, 1, 3, S // 2, S // 2).to(dev); t = torch.full((n,), 700, device=dev)
with torch.no_grad():
    ref = m(x, t, lr, 2).clone()
    plan = m.native_plan(n, n, 1, S, 2)
    eps = torch.empty_like(ref)
    st = N.stream_ptr(dev)
    for i in range(iters):
        N.check(N.lib().drs_unet_forward(plan, N.ptr(x), N.ptr(eps), st))
    torch.cuda.synchronize()
    N.check(N.lib().drs_plan_check(plan, st))
    print("back-to-back ok, max diff vs first:", float((eps - ref).abs().max()))
