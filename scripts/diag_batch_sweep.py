"""Diagnostic: ms per graph-replayed reverse step of the superres 256 x 256 plan by batch size and by whether every
sample has its own condition image (aggregation sampling) or all share one (Diffusion.sample), plus the per-launch
table of one batch size.  usage (GPU box): [DRS_SWEEP=15,16,37,...] python scripts/diag_batch_sweep.py [profile_nb]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import synthetic as T
from diffusionremotesensing_b200 import _native as N
dev = torch.device("cuda:0")
lib = N.lib(); st = N.stream_ptr(dev)
m, _ = T.default_init_model("superres"); m.to(dev).eval()
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=1500, device="cuda:0", magnification_factor=2, image_size=256,
                Degradation_type="DownBlur")
c1, c2, c3 = d._coefficients()
S, K = 256, 30
SIZES = os.environ.get("DRS_SWEEP")
pairs = [(int(v), int(v)) for v in SIZES.split(",")] if SIZES else ((16, 1), (16, 16), (30, 30), (31, 31), (32, 32), (32, 1), (31, 1))
for nb, ncond in pairs:
    plan = m.native_plan(nb, nb, ncond, S, 2)
    cond = T.np_rand(2, ncond, 3, S // 2, S // 2).to(dev)
    x = T.np_randn(3, nb, 3, S, S).to(dev); z = torch.empty_like(x); eps = torch.empty_like(x)
    N.check(lib.drs_cond_encode(plan, N.ptr(cond), st))
    N.check(lib.drs_sampler_prepare(plan, 1500, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))
    N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), 1499, st))
    for _ in range(5):
        z.normal_(); N.check(lib.drs_sampler_step(plan, 1, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        z.normal_(); N.check(lib.drs_sampler_step(plan, 1, st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"nb={nb:3d} ncond={ncond:3d}: {ms:.3f} ms/step = {ms * 1e3 / nb:.1f} us per image-step, {nb * 1e3 / ms:.0f} image-steps/s", flush=True)
    if len(sys.argv) > 1 and nb == int(sys.argv[1]) and ncond == nb:
        nl = lib.drs_plan_launch_count(plan)
        msl = torch.zeros(nl)
        N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 5, N.ptr(msl), st))
        nm = C.create_string_buffer(64)
        ctas, smem = C.c_int(), C.c_int()
        for i in range(nl):
            lib.drs_plan_launch_info(plan, i, nm, 64, None, None, C.byref(ctas), C.byref(smem))
            print(f"   {nm.value.decode():28s} {float(msl[i]) * 1e3:7.1f} us  ctas {ctas.value:6d} smem {smem.value}")
    m.release_native()
