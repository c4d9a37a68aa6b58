"""Diagnostic: GPU time of the parts of one sample_batched() call of aggregation sampling (n patches, K steps):
condition encode, the K reverse steps, the whole call.  usage (GPU box): python scripts/diag_sample_parts.py [n] [K]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import synthetic as T
from diffusionremotesensing_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 121
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
lib = N.lib(); st = N.stream_ptr(dev)
m, _ = T.default_init_model("superres"); m.to(dev).eval()
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=K + 1, device="cuda:0", magnification_factor=2, image_size=256,
                Degradation_type="DownBlur")
lr = T.np_rand(2, n, 3, 128, 128).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
for _ in range(2):
    d.sample_batched(m, lr, generator=gen)
torch.cuda.synchronize()


def ev(fn, reps=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


plan = m.native_plan(n, n, n, 256, 2)
print(f"n = {n}, K = {K}")
print(f"  sample_batched (whole call, events)   {ev(lambda: d.sample_batched(m, lr, generator=gen)):8.2f} ms")
print(f"  drs_cond_encode                       {ev(lambda: N.check(lib.drs_cond_encode(plan, N.ptr(lr), st))):8.2f} ms")
bufs = m.sampler_buffers(n, n, n, 256, 2)
c1, c2, c3 = d._coefficients()
N.check(lib.drs_sampler_prepare(plan, K + 1, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))


def steps():
    N.check(lib.drs_sampler_begin(plan, N.ptr(bufs["x"]), N.ptr(bufs["z"]), N.ptr(bufs["eps"]), K, st))
    for i in range(K, 0, -1):
        if i > 1:
            bufs["z"].normal_(generator=gen)
        N.check(lib.drs_sampler_step(plan, 1, st))


print(f"  {K} reverse steps (normal_ + graph)     {ev(steps):8.2f} ms")
x = torch.empty_like(bufs["x"])
print(f"  randn x_T                             {ev(lambda: x.normal_(generator=gen)):8.2f} ms")
print(f"  two state copies                      {ev(lambda: (bufs['x'].copy_(x), x.copy_(bufs['x']))):8.2f} ms")
t0 = time.perf_counter(); d.sample_batched(m, lr, generator=gen); torch.cuda.synchronize()
print(f"  sample_batched wall clock             {(time.perf_counter() - t0) * 1e3:8.2f} ms")
