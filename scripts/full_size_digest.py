"""Digest of the hot path at BASELINE.json's full size (cfg 2: superres x2, LR 128 -> 256, n = 16): SHA-256 of eps after
one UNet evaluation and of x after three graph-replayed reverse steps with fixed noise. Used by
tests/test_gpu_full_size.py to compare kernel paths selected by environment switches (one process per path)."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import _native as N

n, S = 16, 256
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
lib = N.lib(); st = N.stream_ptr(dev)
lr = T.np_rand(2, 1, 3, S // 2, S // 2).to(dev)
x = T.np_randn(3, n, 3, S, S).to(dev)
with torch.no_grad():
    eps = m(x, torch.full((n,), 700, device=dev), lr, 2)
torch.cuda.synchronize()
print("eps", hashlib.sha256(eps.cpu().numpy().tobytes()).hexdigest())
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=1500, device="cuda:0", magnification_factor=2, image_size=S,
                Degradation_type="DownBlur")
plan = m.native_plan(n, n, 1, S, 2)
c1, c2, c3 = d._coefficients()
z = T.np_randn(4, n, 3, S, S).to(dev); e = torch.empty_like(x)
N.check(lib.drs_cond_encode(plan, N.ptr(lr), st))
N.check(lib.drs_sampler_prepare(plan, 1500, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))
N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(e), 1499, st))
for _ in range(3):
    N.check(lib.drs_sampler_step(plan, 1, st))
N.check(lib.drs_plan_check(plan, st))
torch.cuda.synchronize()
print("x3", hashlib.sha256(x.cpu().numpy().tobytes()).hexdigest(), "finite", bool(torch.isfinite(x).all()))
