"""Diagnostic: the aggregation blend kernel at the cfg-5 shape (961 patches -> 3 x 4096 x 4096), event-timed on a
flushed L2.  usage (GPU box): python scripts/diag_blend.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200.aggregation import blend_patches
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
side, P, stride, k = 2048, 128, 64, 2


class NoDiffusion:
    model = None


agg = D.split_aggregation_sampling(torch.zeros(1, 3, side, side), P, stride, k, NoDiffusion(), "cuda:0")
n = len(agg.patches_lr)
patches = torch.rand((n, 3, P * k, P * k), device=dev)
w2d = agg.weight[0, 0].contiguous()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
H = W = side * k
by = float(patches.numel() * 4 + 3 * H * W * 4 + H * W * 4)
from diffusionremotesensing_b200 import _native as N
coords = torch.tensor([list(i) for i in agg.patches_sr_infos], dtype=torch.int32).contiguous()
out = torch.empty((1, 3, H, W), device=dev); wsum = torch.empty((H, W), device=dev)
for i in range(iters):
    N.check(N.lib().drs_debug_l2_flush(N.ptr(flush), flush.numel(), N.stream_ptr(dev)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    blend_patches(patches, agg.patches_sr_infos, w2d, H, W, clamp=True, out=out, wsum=wsum, coords=coords)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"blend {n} patches: {ms * 1e3:.1f} us, {by / ms / 1e6:.0f} GB/s")
