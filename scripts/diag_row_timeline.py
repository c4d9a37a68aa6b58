"""Diagnostic: timeline of pipeline 0 of CTA 0 of one row-kernel launch of the real cfg-2 plan (16 stamps per row).
usage (GPU box): DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=<launch name> python scripts/diag_row_timeline.py [n] [S]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionremotesensing_b200 import synthetic as T
from diffusionremotesensing_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
x = T.np_randn(1, n, 3, S, S).to(dev)
plan = m.native_plan(n, n, 1, S, 2)
eps = torch.empty_like(x)
st = N.stream_ptr(dev)
lib = N.lib()
nl = lib.drs_plan_launch_count(plan)
ms = torch.zeros(nl)
for _ in range(2):
    N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 3, N.ptr(ms), st))
buf = (C.c_longlong * 512)()
N.check(lib.drs_debug_timeline(buf, 512))
names = ["start", "ring_ok", "afull0", "issued", "prod", "e_arrive", "e_full", "e_done", "runs", "elected", "mmas", "commit", "rowdone"]
t0 = buf[0]
print(os.environ.get("DRS_V2_TIMELINE_LAYER"))
print("row " + " ".join(f"{nm:>8s}" for nm in names))
for r in range(32):
    if buf[r * 16] == 0 and r > 0: continue   # the issuer stamps every second row (two input rows per iteration)
    print(f"{r:3d} " + " ".join(f"{(buf[r * 16 + s] - t0) if buf[r * 16 + s] else -1:8d}" for s in range(13)))
nm = C.create_string_buffer(64)
for i in range(nl):
    lib.drs_plan_launch_info(plan, i, nm, 64, None, None, None, None)
    if os.environ.get("DRS_V2_TIMELINE_LAYER", "?") in nm.value.decode():
        print(f"launch {nm.value.decode()}: event time {float(ms[i]) * 1000:.1f} us")
