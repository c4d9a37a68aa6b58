"""Stress: sampler steps (graph) + profile passes, repeated; reports the first failure (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import _native as N
n, S, reps, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
lib = N.lib(); st = N.stream_ptr(dev)
plan = m.native_plan(n, n, 1, S, 2)
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=1500, device="cuda:0", magnification_factor=2, image_size=S, Degradation_type="DownBlur")
c1, c2, c3 = d._coefficients()
lr = T.np_rand(2, 1, 3, S // 2, S // 2).to(dev)
x = T.np_randn(3, n, 3, S, S).to(dev); z = torch.empty_like(x); eps = torch.empty_like(x)
N.check(lib.drs_cond_encode(plan, N.ptr(lr), st))
N.check(lib.drs_sampler_prepare(plan, 1500, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))
nl = lib.drs_plan_launch_count(plan); ms = torch.zeros(nl)
t0 = time.time()
for r in range(reps):
    try:
        N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), 1499, st))
        for i in range(K):
            z.normal_()
            N.check(lib.drs_sampler_step(plan, 1, st))
        torch.cuda.synchronize()
        N.check(lib.drs_plan_check(plan, st))
        N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), 10, N.ptr(ms), st))
        N.check(lib.drs_plan_check(plan, st))
        x.normal_()
    except Exception as e:
        print("rep", r, "FAILED after", time.time() - t0, "s:", str(e)[:200])
        sys.exit(1)
print("stress ok:", reps, "reps x", K, "steps in", round(time.time() - t0, 1), "s")
