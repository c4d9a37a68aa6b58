"""Diagnostic: aggregation sampling throughput against the patch batch size (run on the GPU box).
usage: python scripts/diag_patch_batch.py [LR scene side] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as T
import diffusionremotesensing_b200 as D

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
m, sd = T.default_init_model("superres"); m.to(dev).eval()
scene = T.np_rand(7, 1, 3, side, side).to(dev)
for pb in (16, 32, 48, 64, 96, 128):
    d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=steps + 1, device=str(dev), magnification_factor=2,
                    image_size=256, Degradation_type="DownBlur")
    agg = D.split_aggregation_sampling(scene, 128, 64, 2, d, str(dev), patch_batch=pb)
    n = len(agg.patches_lr)
    nb = -(-n // pb); size = -(-n // nb)
    agg.sample_patches(range(0, size), private_rng=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = agg.sample_patches(range(n), private_rng=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"patch_batch {pb:4d}: {n} patches in {nb} batches of {size}: {dt * 1e3:8.1f} ms = {dt / (n * steps) * 1e6:6.2f} us per image-step, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del agg, d, out
