#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on ONE GPU (they are parity-test cases, not bench lines;
this script records how fast they run). Each line: config, image-steps/s over K timed reverse steps with CUDA events.

  cfg1  superres x2, LR 32 -> 64, n = 1, cosine 50 steps                 (whole sample() call, host to host)
  cfg3  SAR -> NDVI 128 x 128, n = 4 (one GPU's share of the batch of 32 sharded over 8), cosine 1500
  cfg4  generation 64 x 64, n = 256 / G per GPU with classifier-free guidance (two UNet passes per step), linear 1000
  cfg5  aggregation sampling: LR scene, patch 128 -> 256, stride 64, batched patches (scene size from --scene)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import _native as N

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--scene", type=int, default=512, help="LR scene side for cfg5 (2048 is the BASELINE.json size)")
ap.add_argument("--scene-steps", type=int, default=100, help="noise_steps for cfg5 (1500 is the BASELINE.json value)")
ap.add_argument("--gpus-sim", type=int, default=8, help="cfg4: per-GPU batch = 256 / this")
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = N.lib()
out = []


def timed_steps(model, diff, nb, nx, S, mag, cond, labels, cfg_scale, K):
    st = N.stream_ptr(dev)
    plan = model.native_plan(nb, nx, 1, S, mag)
    c1, c2, c3 = diff._coefficients()
    xc = model._desc().x_channels
    x = torch.randn(nx, xc, S, S, device=dev); z = torch.empty_like(x)
    eps = torch.empty(nb, model._desc().out_channels, S, S, device=dev)
    if cond is not None:
        N.check(lib.drs_cond_encode(plan, N.ptr(cond), st))
    N.check(lib.drs_sampler_prepare(plan, diff.noise_steps, N.ptr(c1), N.ptr(c2), N.ptr(c3), N.ptr(labels), cfg_scale, st))
    N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), diff.noise_steps - 1, st))
    for _ in range(5):
        z.normal_(); N.check(lib.drs_sampler_step(plan, 1, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        z.normal_(); N.check(lib.drs_sampler_step(plan, 1, st))
    e1.record(); torch.cuda.synchronize()
    N.check(lib.drs_plan_check(plan, st))
    return e0.elapsed_time(e1) / K


# cfg1
m, _ = T.default_init_model("superres"); m.to(dev)
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=50, device="cuda:0", magnification_factor=2, image_size=64, Degradation_type="DownBlur")
lr = T.np_rand(1, 3, 32, 32)
d.sample(1, m, lr)
torch.cuda.synchronize(); t0 = time.perf_counter(); r = d.sample(1, m, lr).cpu(); t1 = time.perf_counter()
out.append({"config": "cfg1 superres 32->64 n=1 cosine50, whole sample() host to host", "seconds": t1 - t0, "image_steps_per_s": 49 / (t1 - t0)})
# cfg3
m, _ = T.default_init_model("sar"); m.to(dev).eval()
d = D.Diffusion_SAR_TO_NDVI("cosine", m, "/nonexistent", noise_steps=1500, device="cuda:0", image_size=128)
ms = timed_steps(m, d, 4, 4, 128, 1, T.np_rand(2, 1, 2, 128, 128).to(dev), None, 0.0, args.steps)
out.append({"config": "cfg3 SAR->NDVI 128x128 n=4 per GPU (32 over 8 GPUs) cosine1500", "ms_per_step": ms, "image_steps_per_s": 4e3 / ms})
ms = timed_steps(m, d, 32, 32, 128, 1, T.np_rand(2, 1, 2, 128, 128).to(dev), None, 0.0, args.steps)
out.append({"config": "cfg3 SAR->NDVI 128x128 n=32 on one GPU cosine1500", "ms_per_step": ms, "image_steps_per_s": 32e3 / ms})
# cfg4
m, _ = T.default_init_model("generation"); m.to(dev).eval()
d = D.Diffusion_generation("linear", m, "/nonexistent", noise_steps=1000, device="cuda:0", image_size=64)
for nx in (256 // args.gpus_sim, 256):
    labels = torch.cat([torch.arange(nx, dtype=torch.int32) % 10, torch.full((nx,), -1, dtype=torch.int32)]).contiguous()
    ms = timed_steps(m, d, 2 * nx, nx, 64, 1, None, labels, 3.0, args.steps)
    out.append({"config": f"cfg4 generation 64x64 n={nx} CFG scale 3 (2 passes batched) linear1000", "ms_per_step": ms, "image_steps_per_s": nx * 1e3 / ms})
# cfg5
m, _ = T.default_init_model("superres"); m.to(dev)
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=args.scene_steps, device="cuda:0", magnification_factor=2, image_size=256, Degradation_type="DownBlur")
img = T.np_rand(3, 1, 3, args.scene, args.scene).to(dev)
agg = D.split_aggregation_sampling(img, 128, 64, 2, d, "cuda:0", patch_batch=32)
torch.cuda.synchronize(); t0 = time.perf_counter(); res = agg.aggregation_sampling(); torch.cuda.synchronize(); t1 = time.perf_counter()
n_p = len(agg.patches_lr)
out.append({"config": f"cfg5 aggregation LR {args.scene}^2 -> {2 * args.scene}^2, {n_p} patches 128->256 stride 64, noise_steps {args.scene_steps}, batch 32",
            "seconds": t1 - t0, "image_steps_per_s": n_p * (args.scene_steps - 1) / (t1 - t0),
            "extrapolated_seconds_1499_steps": (t1 - t0) * 1499 / (args.scene_steps - 1)})
for o in out:
    print(json.dumps(o))
