"""Diagnostic: timeline of CTA 0 of one launch of the real cfg-2 plan.
usage (GPU box): DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=<substring of the launch name> python scripts/diag_layer_timeline.py [n] [S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import torch
import common as T
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
t0 = buf[0]
names = ["p_start", "p_issued", "m_tmemfree", "m_afull", "m_issued", "e_tfull", "e_done", "m1_issued"]
print(os.environ.get("DRS_V2_TIMELINE_LAYER"), "total ms", float(ms.sum()))
print("pair " + " ".join(f"{nm:>10s}" for nm in names))
npairs = int(os.environ.get("DRS_TL_PAIRS", "16"))
last = 0
for t in range(min(npairs, 32)):
    if buf[t * 8] == 0 and t > 0: break
    print(f"{t:4d} " + " ".join(f"{buf[t * 8 + s] - t0:10d}" for s in range(8)))
    last = max(last, buf[t * 8 + 6] - t0)
name = os.environ.get("DRS_V2_TIMELINE_LAYER")
import ctypes
nm = ctypes.create_string_buffer(64)
for i in range(nl):
    lib.drs_plan_launch_info(plan, i, nm, 64, None, None, None, None)
    if name and name in nm.value.decode():
        print(f"launch {nm.value.decode()}: event time {float(ms[i]) * 1000:.1f} us; CTA 0 recorded span {last} cycles "
              f"= {last / 1.965e3:.1f} us at 1965 MHz")
if buf[504]:
    print("kernel stamps (cycles from entry): setup done %d, producer past griddep_wait %d, first pair start %d, exit %d" % (
        buf[505] - buf[504], buf[506] - buf[504], buf[0] - buf[504], buf[507] - buf[504]))
tr = [buf[256 + i] for i in range(64)]
if any(tr):
    base = min(v for v in tr if v > 0)
    print("epilogue trace of thread 0, pair 3 (per 16-channel chunk: start, after tmem wait, after math, after store):")
    for c in range(16):
        row = tr[c * 4:c * 4 + 4]
        if any(row):
            print(f"  chunk {c:2d}: " + " ".join(f"{(v - base) if v else -1:7d}" for v in row))
it = [buf[256 + 32 + i] for i in range(30)]
if any(it):
    base = min(v for v in it if v > 0)
    print("issuer 0 trace, pair 1, streamed weights (per ring unit: before the B wait, after it, after MMA issue + commit; -DDRS_EPI_TRACE builds):")
    for k in range(10):
        row = it[k * 3:k * 3 + 3]
        if any(row):
            print(f"  unit {k:2d}: " + " ".join(f"{(v - base) if v else -1:7d}" for v in row))
