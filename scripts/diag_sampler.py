"""Diagnostic: K sampler steps at (n, S) eager or graph; reports finiteness per step (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import _native as N

fam, n, S, K, graph = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
dev = torch.device("cuda:0")
m, sd = T.default_init_model(fam)
m.to(dev).eval()
lib = N.lib(); st = N.stream_ptr(dev)
mag = 2 if fam == "superres" else 1
xc = 3 if fam != "sar" else 1
plan = m.native_plan(n, n, 1, S, mag)
if fam == "superres":
    cond = T.np_rand(2, 1, 3, S // 2, S // 2).to(dev)
elif fam == "sar":
    cond = T.np_rand(2, 1, 2, S, S).to(dev)
else:
    cond = None
d = D.Diffusion_SAR_TO_NDVI("cosine", m, "/nonexistent", noise_steps=1500, device="cuda:0", image_size=S)
c1, c2, c3 = d._coefficients()
x = T.np_randn(3, n, xc, S, S).to(dev); z = torch.empty_like(x); eps = torch.empty_like(x)
if cond is not None:
    N.check(lib.drs_cond_encode(plan, N.ptr(cond), st))
N.check(lib.drs_sampler_prepare(plan, 1500, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))
N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), 1499, st))
bad = None
for i in range(K):
    z.normal_()
    N.check(lib.drs_sampler_step(plan, graph, st))
    if i % 10 == 9 or i == K - 1:
        torch.cuda.synchronize()
        if not bool(torch.isfinite(x).all()) and bad is None:
            bad = i
            break
try:
    N.check(lib.drs_plan_check(plan, st))
    print(fam, n, S, K, "graph" if graph else "eager", "first non-finite step:", bad, "|x|max", float(x.abs().max()))
except Exception as e:
    print("ERR", e)
