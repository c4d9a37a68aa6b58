"""Diagnostic: entry / exit time of every CTA of one row-kernel launch (globaltimer, ns since the first entry).
usage (GPU box): DRS_V2_TIMELINE=8 DRS_V2_TIMELINE_LAYER=<launch name> python scripts/diag_row_cta_times.py"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import synthetic as T
from diffusionremotesensing_b200 import _native as N
n, S = 16, 256
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
x = T.np_randn(1, n, 3, S, S).to(dev)
plan = m.native_plan(n, n, 1, S, 2)
eps = torch.empty_like(x)
st = N.stream_ptr(dev); lib = N.lib()
nl = lib.drs_plan_launch_count(plan)
ms = torch.zeros(nl)
for _ in range(2):
    N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 3, N.ptr(ms), st))
buf = (C.c_longlong * 512)()
N.check(lib.drs_debug_timeline(buf, 512))
ent = [buf[2 * i] for i in range(148)]; ext = [buf[2 * i + 1] for i in range(148)]
t0 = min(e for e in ent if e)
dur = sorted((x1 - e0) / 1e3 for e0, x1 in zip(ent, ext) if e0)
print(os.environ.get("DRS_V2_TIMELINE_LAYER"), "CTAs", len(dur))
print("entry spread us: %.1f .. %.1f" % (min((e - t0) / 1e3 for e in ent if e), max((e - t0) / 1e3 for e in ent if e)))
print("exit  spread us: %.1f .. %.1f" % (min((e - t0) / 1e3 for e in ext if e), max((e - t0) / 1e3 for e in ext if e)))
print("duration us: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (dur[0], dur[len(dur) // 10], dur[len(dur) // 2], dur[9 * len(dur) // 10], dur[-1]))
print("CTA 0: entry %.1f exit %.1f" % ((ent[0] - t0) / 1e3, (ext[0] - t0) / 1e3))
