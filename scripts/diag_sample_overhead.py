"""Diagnostic: where one sample_batched() call of aggregation sampling spends host time (31 patches, K steps)."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import synthetic as T
K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
m, _ = T.default_init_model("superres"); m.to(dev)
d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=K + 1, device="cuda:0", magnification_factor=2, image_size=256,
                Degradation_type="DownBlur")
lr = T.np_rand(2, 31, 3, 128, 128).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
for _ in range(2):
    d.sample_batched(m, lr, generator=gen)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    d.sample_batched(m, lr, generator=gen)
torch.cuda.synchronize()
print(f"sample_batched(31 patches, {K} steps): {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per call")
pr = cProfile.Profile(); pr.enable()
d.sample_batched(m, lr, generator=gen); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
