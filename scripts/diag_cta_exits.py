"""Diagnostic: when does every CTA of one launch finish (load balance of the static tile assignment)?
usage (GPU box): DRS_V2_TIMELINE=12 DRS_V2_TIMELINE_LAYER=<name> DRS_CG2=none python scripts/diag_cta_exits.py"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
from diffusionremotesensing_b200 import _native as N
n, S = 16, 256
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
x = T.np_randn(1, n, 3, S, S).to(dev)
plan = m.native_plan(n, n, 1, S, 2)
eps = torch.empty_like(x)
st = N.stream_ptr(dev)
lib = N.lib()
nl = lib.drs_plan_launch_count(plan)
ms = torch.zeros(nl)
N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 2, N.ptr(ms), st))
N.check(lib.drs_debug_spans(None, 1))
N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 1, N.ptr(ms), st))
buf = (C.c_longlong * 512)()
N.check(lib.drs_debug_timeline(buf, 512))
v = sorted(b / 1e3 for b in buf[:296] if b > 0)
print(os.environ.get("DRS_V2_TIMELINE_LAYER"), "CTAs", len(v))
if v:
    q = lambda f: v[min(len(v) - 1, int(f * len(v)))]
    print(f"exit time after first entry (us): min {v[0]:.1f} p10 {q(.1):.1f} p50 {q(.5):.1f} p90 {q(.9):.1f} max {v[-1]:.1f}")
