"""Diagnostic: drs_plan_profile repeatedly (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
from diffusionremotesensing_b200 import _native as N
n, S, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
x = T.np_randn(1, n, 3, S, S).to(dev); lr = T.np_rand(2, 1, 3, S // 2, S // 2).to(dev); t = torch.full((n,), 700, device=dev)
with torch.no_grad():
    ref = m(x, t, lr, 2).clone()
plan = m.native_plan(n, n, 1, S, 2)
eps = torch.empty_like(ref)
st = N.stream_ptr(dev)
lib = N.lib()
nl = lib.drs_plan_launch_count(plan)
ms = torch.zeros(nl)
for r in range(reps):
    try:
        N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 10, N.ptr(ms), st))
        N.check(lib.drs_plan_check(plan, st))
        print("rep", r, "ok total ms", float(ms.sum()), "diff", float((eps - ref).abs().max()))
    except Exception as e:
        print("rep", r, "FAILED", e)
        break
import ctypes
nm = ctypes.create_string_buffer(64)
print("per-launch ms (last rep):")
for i in range(nl):
    lib.drs_plan_launch_info(plan, i, nm, 64, None, None, None, None)
    print(f"  {nm.value.decode():28s} {float(ms[i]) * 1000:8.1f} us")
