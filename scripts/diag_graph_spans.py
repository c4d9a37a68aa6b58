"""Diagnostic: where one graph-replayed reverse step spends its time. Every tensor-core launch records the
globaltimer at its first CTA's entry and its last CTA's exit (DRS_V2_TIMELINE=4); the table of the last replayed
step gives each launch's in-graph duration and the gap to its predecessor on the same stream.
usage (GPU box): DRS_V2_TIMELINE=4 python scripts/diag_graph_spans.py [n] [S]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
lib = N.lib(); st = N.stream_ptr(dev)
plan = m.native_plan(n, n, 1, S, 2)
cond = T.np_rand(2, 1, 3, S // 2, S // 2).to(dev)
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=1500, device="cuda:0", magnification_factor=2, image_size=S,
                Degradation_type="DownBlur")
c1, c2, c3 = d._coefficients()
x = T.np_randn(3, n, 3, S, S).to(dev); z = torch.empty_like(x); eps = torch.empty_like(x)
N.check(lib.drs_cond_encode(plan, N.ptr(cond), st))
N.check(lib.drs_sampler_prepare(plan, 1500, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))
N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), 1499, st))
for _ in range(20):
    z.normal_(); N.check(lib.drs_sampler_step(plan, 1, st))
torch.cuda.synchronize()
N.check(lib.drs_debug_spans(None, 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); N.check(lib.drs_sampler_step(plan, 1, st)); e1.record()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 128)()
N.check(lib.drs_debug_spans(buf, 0))
nl = lib.drs_plan_launch_count(plan)
nm = C.create_string_buffer(64)
rows = []
for i in range(1, nl):
    lib.drs_plan_launch_info(plan, i, nm, 64, None, None, None, None)
    a, b = buf[2 * (i - 1)], buf[2 * (i - 1) + 1]
    if b > 0 and a != 0xFFFFFFFFFFFFFFFF:
        rows.append((nm.value.decode(), a, b))
t0 = min(r[1] for r in rows)
print(f"step (CUDA events) {e0.elapsed_time(e1) * 1000:.1f} us; first conv entry -> last conv exit {(max(r[2] for r in rows) - t0) / 1e3:.1f} us")
print(f"{'launch':28s} {'start':>8s} {'dur':>7s} {'gap to previous exit on the main chain':>10s}")
prev_end = None
busy = 0.0
for name, a, b in rows:
    gate = name.startswith("gating") or name.startswith("attention")
    gap = "" if (prev_end is None or gate) else f"{(a - prev_end) / 1e3:7.1f}"
    print(f"{name:28s} {(a - t0) / 1e3:8.1f} {(b - a) / 1e3:7.1f} {gap}")
    busy += (b - a) / 1e3
    if not gate:
        prev_end = b
print(f"sum of launch durations {busy:.1f} us")
